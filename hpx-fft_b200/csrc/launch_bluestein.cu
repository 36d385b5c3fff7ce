// launch_bluestein.cu -- plan-side set-up (chirp, transformed chirp filter, work buffers, power-of-two stage) and launchers of the
// Bluestein path (kernels_bluestein.cuh).
#include "kernels_bluestein.cuh"
#include "launch_util.h"

#include <cmath>
#include <complex>

namespace hpxfft_b200 {

struct BlueStage {
    unsigned n = 0, M = 0, strips = 0; // transform length, convolution length (power of two >= 2n-1), strips the buffers hold
    cd *chirp = nullptr;               // c[j] = exp(-i pi j^2 / n), j < n
    cd *hhat = nullptr;                // FFT_M of h[l] = conj(c[|l|]), l in (-n, n), wrapped
    cd *T1 = nullptr, *T2 = nullptr;   // [strips][M][CW] tiled / [M][strips*CW] row-major
    hpxfft_b200_plan *sub = nullptr;   // column stage of length M
};

namespace {

typedef std::complex<long double> lcd;

// in-place iterative radix-2 DIT FFT (forward), long double: only used once per plan for the chirp filter
void host_fft(std::vector<lcd> &a)
{
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    const long double PI_L = 3.14159265358979323846264338327950288L;
    for (size_t len = 2; len <= n; len <<= 1) {
        std::vector<lcd> w(len / 2);
        for (size_t k = 0; k < len / 2; ++k) w[k] = lcd(cosl(2 * PI_L * k / len), -sinl(2 * PI_L * k / len));
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const lcd u = a[i + k], v = a[i + k + len / 2] * w[k];
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}

unsigned grid_for(const hpxfft_b200_plan *p) { return (unsigned) p->sm_count * 8u; }

int blue_core(const hpxfft_b200_plan *p, const BlueStage *b, unsigned strips, int *launches)
{
    // A = FFT_M(a): tiled T1 -> row-major T2 (pitch = strips*CW, what a single-rank column destination produces)
    InterView iv;
    iv.base = b->T1;
    iv.nxl = b->M;
    iv.shift = pow2_shift(iv.nxl);
    iv.tile_stride = (unsigned long long) b->M * CW;
    iv.rank_stride = 0;
    ColDst o;
    o.nxl = b->M;
    o.shift = pow2_shift(o.nxl);
    o.w = strips * CW;
    o.vt = 1;
    o.base[0] = b->T2;
    o.pitch[0] = strips * CW;
    o.col0[0] = 0;
    if (int rc = run_col_stage(b->sub, iv, o, strips, launches)) return rc;
    blue_mul_kernel<<<grid_for(p), BLUE_THREADS, 0, p->stream>>>(b->T2, b->T1, b->hhat, b->M, strips);
    CU(cudaGetLastError());
    if (launches) *launches += 1;
    return run_col_stage(b->sub, iv, o, strips, launches); // p = conj(FFT_M(conj(P))) / M: the conjugations live in the neighbours
}

}  // namespace

int blue_setup(hpxfft_b200_plan *p, bool rows, size_t n, unsigned max_strips)
{
    BlueStage *b = new BlueStage();
    (rows ? p->blue_r : p->blue_c) = b;
    b->n = (unsigned) n;
    b->M = 1;
    while (b->M < 2 * n - 1) b->M <<= 1;
    if (b->M > (1u << 18)) return fail(HPXFFT_B200_EINVAL, "length %zu needs a convolution length above 2^18", n);
    b->strips = max_strips;
    const long double PI_L = 3.14159265358979323846264338327950288L;
    std::vector<double2> c(n), hh(b->M);
    std::vector<lcd> h(b->M, lcd(0, 0));
    for (size_t j = 0; j < n; ++j) {
        const unsigned long long e = ((unsigned long long) j * j) % (2ull * n); // j^2 mod 2n keeps the angle small
        const long double ang = PI_L * (long double) e / (long double) n;
        const lcd cj(cosl(ang), -sinl(ang));
        c[j] = make_double2((double) cj.real(), (double) cj.imag());
        h[j] = std::conj(cj);
        if (j) h[b->M - j] = std::conj(cj);
    }
    host_fft(h);
    for (size_t k = 0; k < b->M; ++k) hh[k] = make_double2((double) h[k].real(), (double) h[k].imag());
    CU(cudaMalloc(&b->chirp, n * sizeof(cd)));
    CU(cudaMemcpy(b->chirp, c.data(), n * sizeof(cd), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&b->hhat, (size_t) b->M * sizeof(cd)));
    CU(cudaMemcpy(b->hhat, hh.data(), (size_t) b->M * sizeof(cd), cudaMemcpyHostToDevice));
    const size_t bytes = (size_t) max_strips * b->M * CW * sizeof(cd);
    CU(cudaMalloc(&b->T1, bytes));
    CU(cudaMalloc(&b->T2, bytes));
    return make_col_stage(p, b->M, max_strips, &b->sub);
}

void blue_free(hpxfft_b200_plan *p)
{
    for (BlueStage **pb : {&p->blue_r, &p->blue_c}) {
        BlueStage *b = *pb;
        if (!b) continue;
        cudaFree(b->chirp);
        cudaFree(b->hhat);
        cudaFree(b->T1);
        cudaFree(b->T2);
        free_col_stage(b->sub);
        delete b;
        *pb = nullptr;
    }
}

int launch_cols_blue(const hpxfft_b200_plan *p, const InterView &in, const ColDst &out, int *launches)
{
    const BlueStage *b = p->blue_c;
    const unsigned strips = p->ntiles;
    blue_cols_pre_kernel<<<grid_for(p), BLUE_THREADS, 0, p->stream>>>(in, b->T1, b->chirp, b->n, b->M, strips);
    CU(cudaGetLastError());
    if (int rc = blue_core(p, b, strips, launches)) return rc;
    blue_cols_post_kernel<<<grid_for(p), BLUE_THREADS, 0, p->stream>>>(b->T2, out, b->chirp, b->n, b->M, strips);
    CU(cudaGetLastError());
    if (launches) *launches += 2;
    return 0;
}

int launch_rows_blue(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch)
{
    const BlueStage *b = p->blue_r;
    const unsigned strips = (nrows + CW - 1) / CW;
    if (strips > b->strips) return fail(HPXFFT_B200_ESTATE, "Bluestein row buffers hold %u strips, %u needed", b->strips, strips);
    blue_rows_pre_kernel<<<grid_for(p), BLUE_THREADS, 0, p->stream>>>(V, pitch, nrows, b->T1, b->chirp, b->n, b->M, strips);
    CU(cudaGetLastError());
    if (int rc = blue_core(p, b, strips, nullptr)) return rc;
    blue_rows_post_kernel<<<grid_for(p), BLUE_THREADS, 0, p->stream>>>(b->T2, dst, b->chirp, p->tw_row, b->n, b->M, strips, nrows);
    CU(cudaGetLastError());
    return 0;
}

}  // namespace hpxfft_b200

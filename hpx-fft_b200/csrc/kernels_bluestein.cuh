// kernels_bluestein.cuh -- lengths with a large prime factor (no mixed-radix plan): Bluestein's chirp-z algorithm on top of the
// power-of-two column kernels.  FFTW accepts every length (core/src/util/adapter_fftw.cpp:6-10,24-30); this is the
// catch-all behind the Stockham and mixed-radix paths, correct for any n but five passes over memory per dimension.
//
//   X[k] = c[k] * sum_j (x[j] c[j]) conj(c[k-j]),   c[j] = exp(-i pi j^2 / n)
// i.e. a convolution of a[j] = x[j] c[j] with h[l] = conj(c[l]), evaluated as a circular convolution of length
// M = 2^p >= 2n - 1:  A = FFT_M(a),  P = A .* H (H = FFT_M(h), precomputed),  p = IFFT_M(P) = conj(FFT_M(conj(P))) / M.
// Both transforms run on 16-sequence tiles [strip][index][c] with the ordinary column kernels (a plan-view of length M);
// for rows a strip is 16 adjacent rows, transposed on the way in and out.
#pragma once
#include "kernels_rows.cuh"

namespace hpxfft_b200 {

constexpr int BLUE_THREADS = 256;

// columns, step 1: T1[strip][x][c] = I(x, strip, c) * c[x] for x < n, 0 for n <= x < M
__global__ void blue_cols_pre_kernel(InterView in, cd *__restrict__ T1, const cd *__restrict__ chirp, unsigned n, unsigned M, unsigned strips)
{
    const unsigned long long total = (unsigned long long) strips * M * CW;
    for (unsigned long long idx = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned c = (unsigned) (idx % CW);
        const unsigned long long t = idx / CW;
        const unsigned x = (unsigned) (t % M), s = (unsigned) (t / M);
        T1[idx] = x < n ? cmul(ld_stream(inter_ptr(in, x, s, c)), ldtw(chirp, x)) : make_double2(0.0, 0.0);
    }
}

// rows, step 1: strip s = rows 16 s .. 16 s + 15;  T1[s][j][c] = z[16 s + c][j] * c[j] for j < m, 0 otherwise (a transposing read)
__global__ void blue_rows_pre_kernel(const cd *__restrict__ V, unsigned pitch, unsigned nrows, cd *__restrict__ T1, const cd *__restrict__ chirp,
                                     unsigned m, unsigned M, unsigned strips)
{
    const unsigned long long total = (unsigned long long) strips * M * CW;
    for (unsigned long long idx = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned c = (unsigned) (idx % CW);
        const unsigned long long t = idx / CW;
        const unsigned j = (unsigned) (t % M), s = (unsigned) (t / M);
        const unsigned row = s * CW + c;
        T1[idx] = (j < m && row < nrows) ? cmul(V[(unsigned long long) row * pitch + j], ldtw(chirp, j)) : make_double2(0.0, 0.0);
    }
}

// step 3: T1[strip][k][c] = conj(T2[k][strip*CW + c] * H[k])   (T2 is row-major with pitch P = strips*CW, as the column kernels leave it)
__global__ void blue_mul_kernel(const cd *__restrict__ T2, cd *__restrict__ T1, const cd *__restrict__ hhat, unsigned M, unsigned strips)
{
    const unsigned long long total = (unsigned long long) strips * M * CW, P = (unsigned long long) strips * CW;
    for (unsigned long long idx = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned c = (unsigned) (idx % CW);
        const unsigned long long t = idx / CW;
        const unsigned k = (unsigned) (t % M), s = (unsigned) (t / M);
        T1[idx] = cconj(cmul(T2[k * P + (unsigned long long) s * CW + c], ldtw(hhat, k)));
    }
}

// p[k] of sequence (strip s, lane c) from the second transform's output
__device__ __forceinline__ cd blue_result(const cd *__restrict__ T2, unsigned long long P, unsigned s, unsigned c, unsigned k, double invM)
{
    const cd v = T2[k * P + (unsigned long long) s * CW + c];
    return make_double2(v.x * invM, -v.y * invM);
}

// columns, step 5: X[k] = c[k] p[k], k < n, written through the ordinary column destination
__global__ void blue_cols_post_kernel(const cd *__restrict__ T2, ColDst out, const cd *__restrict__ chirp, unsigned n, unsigned M, unsigned strips)
{
    const unsigned long long total = (unsigned long long) strips * n * CW, P = (unsigned long long) strips * CW;
    const double invM = 1.0 / (double) M;
    for (unsigned long long idx = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned c = (unsigned) (idx % CW);
        const unsigned long long t = idx / CW;
        const unsigned k = (unsigned) (t % n), s = (unsigned) (t / n);
        const unsigned kl = s * CW + c;
        if (kl < out.w) st_stream(coldst_ptr(out, k, kl), cmul(blue_result(T2, P, s, c, k, invM), ldtw(chirp, k)));
    }
}

// rows, step 5: Z[k] = c[k] p[k] (half-length complex spectrum), then the Hermitian split of the r2c transform, k <= m/2
__global__ void blue_rows_post_kernel(const cd *__restrict__ T2, RowDst dst, const cd *__restrict__ chirp, const cd *__restrict__ tw, unsigned m, unsigned M,
                                      unsigned strips, unsigned nrows)
{
    const unsigned half = m / 2 + 1; // k = 0 .. floor(m/2)
    const unsigned long long total = (unsigned long long) strips * half * CW, P = (unsigned long long) strips * CW;
    const double invM = 1.0 / (double) M;
    for (unsigned long long idx = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned c = (unsigned) (idx % CW);
        const unsigned long long t = idx / CW;
        const unsigned k = (unsigned) (t % half), s = (unsigned) (t / half);
        const unsigned row = s * CW + c;
        if (row >= nrows) continue;
        auto Z = [&](unsigned q) -> cd { return cmul(blue_result(T2, P, s, c, q, invM), ldtw(chirp, q)); };
        if (k == 0) {
            const cd z0 = Z(0);
            *rowdst_ptr(dst, row, 0u) = make_double2(z0.x + z0.y, 0.0);
            *rowdst_ptr(dst, row, m) = make_double2(z0.x - z0.y, 0.0);
        } else if (2 * k == m) {
            *rowdst_ptr(dst, row, k) = cconj(Z(k));
        } else {
            cd xk, xmk;
            herm_pair(Z(k), Z(m - k), ldtw(tw, k), xk, xmk);
            *rowdst_ptr(dst, row, k) = xk;
            *rowdst_ptr(dst, row, m - k) = xmk;
        }
    }
}

}  // namespace hpxfft_b200

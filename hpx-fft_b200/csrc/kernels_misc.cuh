// kernels_misc.cuh -- synthetic input fill, layout conversion helpers, exchange#2 unpack.
#pragma once
#include "layout.cuh"

namespace hpxfft_b200 {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ double u64_to_unit(unsigned long long x)
{
    return (double) (x >> 11) * (1.0 / 4503599627370496.0) - 1.0; // top 53 bits -> [-1, 1)
}

// Same definitions as oracle/oracle.py:make_input.  row0 = global index of local row 0.
__global__ void fill_kernel(double *__restrict__ V, unsigned nxl, unsigned ny, unsigned n_col, unsigned long long row0,
                            int pattern, unsigned long long seed)
{
    const unsigned long long total = (unsigned long long) nxl * n_col;
    for (unsigned long long idx = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned i = (unsigned) (idx / n_col), j = (unsigned) (idx % n_col);
        double val = 0.0;
        if (j < ny) {
            const unsigned long long ig = row0 + i;
            if (pattern == 0) {
                val = (double) j;
            } else if (pattern == 1) {
                val = u64_to_unit(splitmix64((ig * ny + j) ^ seed));
            } else {
                for (int r = 0; r < 4; ++r) {
                    const double a = u64_to_unit(splitmix64(ig ^ (seed + 1000ull + r)));
                    const double b = u64_to_unit(splitmix64((unsigned long long) j ^ (seed + 2000ull + r)));
                    val += a * b;
                }
            }
        }
        V[idx] = val;
    }
}

// exchange#2 unpack: recv2 holds, for every source rank q, a dense [nxl][w_q] block; scatter it into
// V[j][c_q + kl].  Replaces transpose_x_to_y (core/src/distributed/loop.cpp:108-127) in natural ky order.
// grid = (nxl, P)
__global__ void unpack_kernel(const cd *__restrict__ recv2, cd *__restrict__ V, unsigned nxl, unsigned cy, unsigned wq0, unsigned P,
                              unsigned me)
{
    const unsigned j = blockIdx.x, q = blockIdx.y;
    if (q == me) return; // this rank's own columns were written straight into V by the column kernel
    const unsigned c0 = q * wq0;
    const unsigned w = (q == P - 1) ? cy - c0 : wq0;
    const cd *src = recv2 + (unsigned long long) nxl * c0 + (unsigned long long) j * w;
    cd *dst = V + (unsigned long long) j * cy + c0;
    for (unsigned k = threadIdx.x; k < w; k += blockDim.x) st_stream(dst + k, ld_stream(src + k));
}

// chunked variant for the pipelined exchange #2 (UnpackChunk: layout.cuh)
// grid = (nxl, P)
__global__ void unpack_chunk_kernel(const cd *__restrict__ recv, cd *__restrict__ V, unsigned cy, UnpackChunk u)
{
    const unsigned j = blockIdx.x, q = blockIdx.y;
    const unsigned w = u.wc[q];
    if (w == 0) return;
    const cd *src = recv + u.src_off[q] + (unsigned long long) j * w;
    cd *dst = V + (unsigned long long) j * cy + u.dst_col[q];
    for (unsigned k = threadIdx.x; k < w; k += blockDim.x) st_stream(dst + k, ld_stream(src + k));
}

// ---- generic lengths (any even ny, any nx): direct O(n^2) DFT on the GPU ---------------------------
// FFTW accepts every length and the reference's default example is 8 x 14 (examples/hpxfft/
// shared_loop_2d.cpp:142-143).  Sizes that are not powers of two take these kernels -- slow but exact to
// ~sqrt(n) ulp, and still no CPU fallback.  tw = w_n^i, i < n.
// grid = (nxl, ceil(cy/128)): one thread per output bin of one row
__global__ void rows_generic_kernel(const double *__restrict__ V, unsigned n_col, unsigned nxl, unsigned ny, RowDst dst,
                                    const cd *__restrict__ tw)
{
    const unsigned row = blockIdx.x, k = blockIdx.y * blockDim.x + threadIdx.x;
    if (row >= nxl || k > ny / 2) return;
    const double *x = V + (unsigned long long) row * n_col;
    double re = 0.0, im = 0.0;
    unsigned idx = 0; // (j * k) mod ny, updated incrementally
    for (unsigned j = 0; j < ny; ++j) {
        const cd w = ldtw(tw, idx);
        re += x[j] * w.x;
        im += x[j] * w.y;
        idx += k;
        if (idx >= ny) idx -= ny;
    }
    if (k == 0 || 2 * k == ny) im = 0.0;
    *rowdst_ptr(dst, row, k) = make_double2(re, im);
}
// grid = (ntiles * CW, ceil(nx/128)): one thread per (kx, local column)
__global__ void cols_generic_kernel(InterView in, ColDst out, unsigned nx, const cd *__restrict__ tw)
{
    const unsigned kx = blockIdx.y * blockDim.x + threadIdx.x, kl = blockIdx.x;
    if (kx >= nx || kl >= out.w) return;
    const unsigned ct = kl / CW, c = kl % CW;
    double re = 0.0, im = 0.0;
    unsigned idx = 0;
    for (unsigned x = 0; x < nx; ++x) {
        const cd y = *inter_ptr(in, x, ct, c);
        const cd w = ldtw(tw, idx);
        re += y.x * w.x - y.y * w.y;
        im += y.x * w.y + y.y * w.x;
        idx += kx;
        if (idx >= nx) idx -= nx;
    }
    *coldst_ptr(out, kx, kl) = make_double2(re, im);
}

// row-major [n][width] complex -> column-tiled I[ct][x][c] (one rank); test / adapter entry points only
__global__ void tile_kernel(const cd *__restrict__ A, cd *__restrict__ I, unsigned n, unsigned width)
{
    const unsigned ntiles = (width + CW - 1) / CW;
    const unsigned long long total = (unsigned long long) ntiles * n * CW;
    for (unsigned long long idx = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned c = (unsigned) (idx % CW);
        const unsigned long long t = idx / CW;
        const unsigned x = (unsigned) (t % n), ct = (unsigned) (t / n);
        const unsigned k = ct * CW + c;
        I[idx] = k < width ? A[(unsigned long long) x * width + k] : make_double2(0.0, 0.0);
    }
}
// inverse of tile_kernel
__global__ void untile_kernel(const cd *__restrict__ I, cd *__restrict__ A, unsigned n, unsigned width)
{
    const unsigned long long total = (unsigned long long) n * width;
    for (unsigned long long idx = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned k = (unsigned) (idx % width), x = (unsigned) (idx / width);
        A[idx] = I[((unsigned long long) (k / CW) * n + x) * CW + (k % CW)];
    }
}

}  // namespace hpxfft_b200

// kernels_cols.cuh -- forward c2c FFT along x (the strided axis) on column tiles.
//
// Replaces fft_1d_c2c_inplace on the transposed array plus both local transposes of the reference
// (core/src/shared/loop.cpp:11-15,18-25,46-53; core/src/distributed/loop.cpp:12-16,87-127):
// instead of transposing so that x becomes contiguous, a CTA owns a tile of CW = 16 adjacent ky
// columns and runs the FFT down the rows.  All global accesses are 256-byte segments, the shared
// memory tile is [point][column] with the column index fastest across threads, which makes every
// shared-memory access conflict-free without padding.
//
// nx <= 256        : one Stockham FFT per tile (cols_single_kernel).
// nx = n1*n2 > 256 : four-step.  Level A (cols_levelA_kernel): for every x2, FFT over x1 (stride n2
//                    rows), multiply by w_nx^(k1*x2), write scratch S[ct][k1][x2][c].  Level B
//                    (cols_levelB_kernel): for every k1, FFT over x2 (contiguous in S), result row is
//                    kx = k1 + n1*k2, written straight to its final position (ColDst).
#pragma once
#include "layout.cuh"

namespace hpxfft_b200 {

__host__ __device__ constexpr int col_npass(int N) { return N <= 16 ? 1 : (N <= 256 ? 2 : 3); }
__host__ __device__ constexpr int col_radix(int N, int p)
{
    if (N <= 16) return N;
    if (N == 32) return p == 0 ? 8 : 4;
    if (N == 64) return 8;
    if (N == 128) return p == 0 ? 16 : 8;
    if (N == 256) return 16;
    return 8; // 512 = 8*8*8
}
__host__ __device__ constexpr int col_pt(int N) { return N < 16 ? N : 16; }       // points per thread
__host__ __device__ constexpr int col_threads(int N) { return (N / col_pt(N)) * CW; }
__host__ __device__ constexpr size_t col_smem_bytes(int N) { return N <= 16 ? 0 : (size_t) N * CW * sizeof(cd); }

template <int N, int PT, int R, int NS, bool FIRST, bool LAST, class LD, class ST>
__device__ __forceinline__ void col_pass(cd (&v)[PT], cd *smem, const cd *__restrict__ tw, unsigned tws, LD &ld, ST &st,
                                         int u, int c)
{
    constexpr int NB = PT / R, T = N / R, U = N / PT;
    constexpr int LGR = ilog2(R);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int j = u + b * U;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (FIRST)
                v[b * R + r] = ld(j + r * T, c);
            else
                v[b * R + r] = smem[(j + r * T) * CW + c];
        }
    }
    if (!FIRST && !LAST) __syncthreads(); // every thread has read its inputs before the in-place overwrite
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int j = u + b * U;
        const int k = j & (NS - 1);
        cd w[R];
#pragma unroll
        for (int r = 0; r < R; ++r) w[r] = v[b * R + r];
        if (NS > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) w[r] = cmul(w[r], ldtw(tw, (unsigned) (r * k * (N / (NS * R))) * tws));
        }
        fft_dif<R>(w);
        const int j0 = ((j - k) << LGR) + k; // (j / NS) * NS * R + k
#pragma unroll
        for (int s = 0; s < R; ++s) {
            const int o = j0 + s * NS;
            if (LAST)
                st(o, c, w[bitrev(s, LGR)]);
            else
                smem[o * CW + c] = w[bitrev(s, LGR)];
        }
    }
    if (!LAST) __syncthreads();
}

// One length-N forward FFT down each of the CW columns of a tile.  blockDim.x == col_threads(N).
template <int N, class LD, class ST>
__device__ __forceinline__ void tile_fft(cd *smem, const cd *__restrict__ tw, unsigned tws, LD &ld, ST &st)
{
    constexpr int PT = col_pt(N);
    constexpr int NP = col_npass(N);
    constexpr int R0 = col_radix(N, 0);
    const int c = threadIdx.x % CW, u = threadIdx.x / CW;
    cd v[PT];
    if constexpr (NP == 1) {
        col_pass<N, PT, R0, 1, true, true>(v, smem, tw, tws, ld, st, u, c);
    } else if constexpr (NP == 2) {
        constexpr int R1 = col_radix(N, 1);
        col_pass<N, PT, R0, 1, true, false>(v, smem, tw, tws, ld, st, u, c);
        col_pass<N, PT, R1, R0, false, true>(v, smem, tw, tws, ld, st, u, c);
    } else {
        constexpr int R1 = col_radix(N, 1), R2 = col_radix(N, 2);
        col_pass<N, PT, R0, 1, true, false>(v, smem, tw, tws, ld, st, u, c);
        col_pass<N, PT, R1, R0, false, false>(v, smem, tw, tws, ld, st, u, c);
        col_pass<N, PT, R2, R0 * R1, false, true>(v, smem, tw, tws, ld, st, u, c);
    }
}

// ---- nx <= 256: whole column in one tile -------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(col_threads(N)) cols_single_kernel(InterView in, ColDst out, const cd *__restrict__ tw)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *smem = reinterpret_cast<cd *>(smem_raw);
    const unsigned ct = blockIdx.x;
    auto ld = [&](int i, int c) -> cd { return ld_stream(inter_ptr(in, (unsigned) i, ct, (unsigned) c)); };
    auto st = [&](int k, int c, cd val) {
        const unsigned kl = ct * CW + c;
        if (kl < out.w) st_stream(coldst_ptr(out, (unsigned) k, kl), val);
    };
    tile_fft<N>(smem, tw, 1u, ld, st);
}

// ---- four-step level A: FFT over x1 for fixed x2, then inter-level twiddle ----------------------
template <int N1>
__global__ void __launch_bounds__(col_threads(N1))
    cols_levelA_kernel(InterView in, cd *__restrict__ S, unsigned n2, const cd *__restrict__ tw)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *smem = reinterpret_cast<cd *>(smem_raw);
    const unsigned x2 = blockIdx.x, ct = blockIdx.y;
    cd *Sct = S + (unsigned long long) ct * N1 * n2 * CW;
    auto ld = [&](int i, int c) -> cd { return ld_stream(inter_ptr(in, (unsigned) i * n2 + x2, ct, (unsigned) c)); };
    auto st = [&](int k1, int c, cd val) {
        const cd w = ldtw(tw, (unsigned) k1 * x2); // w_nx^(k1*x2)
        Sct[((unsigned long long) k1 * n2 + x2) * CW + c] = cmul(val, w);
    };
    tile_fft<N1>(smem, tw, n2, ld, st);
}

// ---- four-step level B: FFT over x2 for fixed k1, output row kx = k1 + n1*k2 ---------------------
template <int N2>
__global__ void __launch_bounds__(col_threads(N2))
    cols_levelB_kernel(const cd *__restrict__ S, ColDst out, unsigned n1, const cd *__restrict__ tw)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *smem = reinterpret_cast<cd *>(smem_raw);
    const unsigned k1 = blockIdx.x, ct = blockIdx.y;
    const cd *Sk = S + ((unsigned long long) ct * n1 + k1) * N2 * CW;
    auto ld = [&](int i, int c) -> cd { return Sk[(unsigned) i * CW + c]; };
    auto st = [&](int k2, int c, cd val) {
        const unsigned kl = ct * CW + c;
        if (kl < out.w) st_stream(coldst_ptr(out, k1 + n1 * (unsigned) k2, kl), val);
    };
    tile_fft<N2>(smem, tw, n1, ld, st);
}

}  // namespace hpxfft_b200

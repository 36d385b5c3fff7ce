// kernels_rows16.cuh -- EXPERIMENTAL 512-thread variant of the row r2c kernel for ny = 16384 (m = 8192).
//
// Status: compiles for sm_100a and its index algebra is checked on the CPU (tools/model_kernels.py,
// tests/test_kernel_model.py::test_paired_radix8_last_pass and the m = 8192 run quoted in DESIGN.md), but it
// has NOT been run on hardware yet (the round-1 GPU budget was spent).  It is therefore opt-in only:
// HPXFFT_B200_ROWS16=1 selects it, the default path is rows_r2c_kernel<8192,1,*> (kernels_rows.cuh).
//
// Why: the 256-thread kernel keeps 8 warps per SM (one 128 KB pencil per SM, 244 registers per thread) and is
// latency bound (issue 25 %, FP64 pipe 33 %).  This variant runs the same row with 512 threads x 16 points
// (<= 128 registers), i.e. 16 warps per SM, at the price of one more shared-memory pass:
//   Stockham plan 16 x 16 x 4, then a paired radix-8 last pass (columns j and PP-j, PP = m/8 = 1024) fused with
//   the Hermitian split exactly like the radix-16 version (conjugate twiddles + one-slot output rotation for
//   the mirrored column; split twiddle = w_n^j times a compile-time w_16^s).
// The next row is staged into the dead pencil buffer with cp.async as in the 256-thread kernel.
#pragma once
#include "kernels_rows.cuh"

namespace hpxfft_b200 {

namespace rows16 {
constexpr int M = 8192, T = 512, PT = 16, PS = 4;
constexpr int PP = M / 8;       // columns of the last pass
constexpr int JW = PP / 2 + 1;  // 513
constexpr int LP = M + (M >> PS);
// shared memory: pencil | tw1[16][16] (pass 2) | twb[3][256] (pass 3) | tw2[7][JW] (last pass) | tw3[JW] (split)
constexpr int TW1 = 256, TWB = 3 * 256, TW2 = 7 * JW, TW3 = JW;
constexpr size_t SMEM = (size_t) (LP + TW1 + TWB + TW2 + TW3) * sizeof(cd);
static_assert(SMEM <= 227 * 1024, "rows16 shared-memory budget");

__device__ __forceinline__ void stage(cd *sm, const cd *__restrict__ zrow, int lt)
{
#pragma unroll
    for (int e = 0; e < PT; ++e) cp_async16(sm + lt + e * T, zrow + lt + e * T);
}
}  // namespace rows16

// tw: w_n^i, i < n = 2*8192.  grid = persistent CTAs (one per SM).
template <bool FASTADDR>
__global__ void __launch_bounds__(rows16::T, 1)
    rows16_r2c_kernel(const cd *__restrict__ V, unsigned pitch, unsigned nxl, RowDst dst, const cd *__restrict__ tw)
{
    using namespace rows16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *sm = reinterpret_cast<cd *>(smem_raw);
    cd *tw1 = sm + LP;
    cd *twb = tw1 + TW1;
    cd *tw2 = twb + TWB;
    cd *tw3 = tw2 + TW2;
    const int lt = threadIdx.x;

    // twiddle tables (published by the first barrier of the row loop)
    for (int i = lt; i < TW1; i += T) {
        const int r = i >> 4, k = i & 15;                       // w_256^(r k) = w_n^(64 r k)
        tw1[i] = ldtw(tw, (unsigned) (r * k) * 64u);
    }
    for (int i = lt; i < TWB; i += T) {
        const int r = i / 256 + 1, k = i & 255;                 // w_1024^(r k) = w_n^(16 r k), r = 1..3
        twb[i] = ldtw(tw, (unsigned) (r * k) * 16u);
    }
    for (int i = lt; i < TW2; i += T) {
        const int r = i / JW + 1, j = i - (r - 1) * JW;          // w_M^(r j) = w_n^(2 r j), r = 1..7, j <= PP/2
        tw2[i] = ldtw(tw, 2u * (unsigned) (r * j));
    }
    for (int i = lt; i < TW3; i += T) tw3[i] = ldtw(tw, (unsigned) i); // w_n^j

    auto row_ptr = [&](unsigned row) -> const cd * { return V + (unsigned long long) (row < nxl ? row : nxl - 1) * pitch; };
    if (blockIdx.x < nxl) stage(sm, row_ptr(blockIdx.x), lt);

    for (unsigned row = blockIdx.x; row < nxl; row += gridDim.x) {
        cd v[PT];
        // ---- pass 1: radix 16, NS = 1, inputs = staged raw row (identity layout, thread-private) ----
        cp_async_wait_all();
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = sm[lt + r * (M / 16)];
        __syncthreads();
        fft_dif<16>(v);
#pragma unroll
        for (int s = 0; s < 16; ++s) sm[rpad<PS>(lt * 16 + s)] = v[bitrev(s, 4)];
        __syncthreads();
        // ---- pass 2: radix 16, NS = 16 ----
        {
            const int j = lt, k = j & 15;
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r] = sm[rpad<PS>(j + r * (M / 16))];
            __syncthreads();
#pragma unroll
            for (int r = 1; r < 16; ++r) v[r] = cmul(v[r], tw1[r * 16 + k]);
            fft_dif<16>(v);
            const int j0 = ((j - k) << 4) + k;
#pragma unroll
            for (int s = 0; s < 16; ++s) sm[rpad<PS>(j0 + s * 16)] = v[bitrev(s, 4)];
            __syncthreads();
        }
        // ---- pass 3: radix 4, NS = 256, four butterflies per thread ----
        {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int j = lt + b * T;
#pragma unroll
                for (int r = 0; r < 4; ++r) v[b * 4 + r] = sm[rpad<PS>(j + r * (M / 4))];
            }
            __syncthreads();
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int j = lt + b * T, k = j & 255;
                cd w[4];
                w[0] = v[b * 4];
#pragma unroll
                for (int r = 1; r < 4; ++r) w[r] = cmul(v[b * 4 + r], twb[(r - 1) * 256 + k]);
                fft_dif<4>(w);
                const int j0 = ((j - k) << 2) + k;
#pragma unroll
                for (int s = 0; s < 4; ++s) sm[rpad<PS>(j0 + s * 256)] = w[bitrev(s, 2)];
            }
            __syncthreads();
        }
        // ---- last pass: two radix-8 butterflies (columns jA, jB = PP - jA) + Hermitian split ----
        const int jA = lt, jB = lt ? PP - lt : PP / 2;
        cd A[8], B[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            A[r] = sm[rpad<PS>(jA + r * PP)];
            B[r] = sm[rpad<PS>(jB + r * PP)];
        }
        __syncthreads(); // pencil buffer dead: refill it with the next row while this one finishes in registers
        if (row + gridDim.x < nxl) stage(sm, row_ptr(row + gridDim.x), lt);
        if (lt != 0) {
#pragma unroll
            for (int r = 1; r < 8; ++r) {
                const cd t = tw2[(r - 1) * JW + jA];
                A[r] = cmul(A[r], t);
                B[r] = cmulc(B[r], t); // w_M^(r (PP - jA)) = w_8^r conj(w_M^(r jA)); w_8^r rotates the output by one
            }
        } else {
#pragma unroll
            for (int r = 1; r < 8; ++r) B[r] = mulw32(B[r], 2 * r); // jA = 0; jB = PP/2: w_M^(r PP/2) = w_16^r
        }
        fft_dif<8>(A);
        fft_dif<8>(B);
        if (lt != 0) {
            // Z[jA + s PP] = A[bitrev(s)];  Z[M - (jA + s PP)] = natural output 7-s of column jB = B[bitrev((8-s)&7)]
            const cd wb = tw3[jA];
            cd *pk = nullptr, *pm = nullptr;
            long long step = 0;
            if constexpr (FASTADDR) {
                const unsigned k0 = (unsigned) jA, m0 = (unsigned) M - k0;
                pk = dst.base[0] + (unsigned long long) (k0 >> 4) * dst.tile_stride + (unsigned long long) row * CW + (k0 & 15u);
                pm = dst.base[0] + (unsigned long long) (m0 >> 4) * dst.tile_stride + (unsigned long long) row * CW + (m0 & 15u);
                step = (long long) (PP >> 4) * (long long) dst.tile_stride;
            }
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                cd xk, xmk;
                herm_pair(A[bitrev(s, 3)], B[bitrev((8 - s) & 7, 3)], mulw32(wb, 2 * s), xk, xmk); // w_n^(jA + s PP) = w_n^jA w_16^s
                if constexpr (FASTADDR) {
                    st_stream(pk + s * step, xk);
                    st_stream(pm - s * step, xmk);
                } else {
                    const unsigned kA = (unsigned) (jA + s * PP);
                    st_stream(rowdst_ptr(dst, row, kA), xk);
                    st_stream(rowdst_ptr(dst, row, (unsigned) M - kA), xmk);
                }
            }
        } else {
            // columns 0 and PP/2 are their own partners
            const cd z0 = A[0];
            st_stream(rowdst_ptr(dst, row, 0u), make_double2(z0.x + z0.y, 0.0));
            st_stream(rowdst_ptr(dst, row, (unsigned) M), make_double2(z0.x - z0.y, 0.0));
#pragma unroll
            for (int s = 1; s < 4; ++s) {
                cd xk, xmk;
                herm_pair(A[bitrev(s, 3)], A[bitrev(8 - s, 3)], mulw32(make_double2(1.0, 0.0), 2 * s), xk, xmk); // w_n^(s PP) = w_16^s
                st_stream(rowdst_ptr(dst, row, (unsigned) (s * PP)), xk);
                st_stream(rowdst_ptr(dst, row, (unsigned) (M - s * PP)), xmk);
            }
            st_stream(rowdst_ptr(dst, row, (unsigned) (4 * PP)), cconj(A[bitrev(4, 3)]));
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                cd xk, xmk;
                // k = PP/2 + s PP: w_n^k = w_32^(1 + 2 s); partner = natural output 7-s of the same column
                herm_pair(B[bitrev(s, 3)], B[bitrev(7 - s, 3)], mulw32(make_double2(1.0, 0.0), 1 + 2 * s), xk, xmk);
                st_stream(rowdst_ptr(dst, row, (unsigned) (PP / 2 + s * PP)), xk);
                st_stream(rowdst_ptr(dst, row, (unsigned) (M - PP / 2 - s * PP)), xmk);
            }
        }
    }
}

}  // namespace hpxfft_b200

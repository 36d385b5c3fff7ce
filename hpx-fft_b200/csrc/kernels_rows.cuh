// kernels_rows.cuh -- in-place-semantics 1-D r2c FFT of every local row (length ny, ny even).
//
// Replaces fft_1d_r2c_inplace -> fftw_execute_dft_r2c (core/src/shared/loop.cpp:6-9,
// core/src/distributed/loop.cpp:7-10, core/src/util/adapter_fftw.cpp:12-15) fused with the pack /
// transpose that follows it in the reference (shared/loop.cpp:18-25, distributed/loop.cpp:19-27):
// the result is written directly in the column-tiled intermediate layout (layout.cuh).
//
// Algorithm: z[j] = x[2j] + i x[2j+1] (a row of ny doubles *is* ny/2 double2), length-m complex
// Stockham FFT (m = ny/2) with the whole pencil resident in shared memory; the raw row is staged into
// the pencil buffer by cp.async while the PREVIOUS row finishes in registers, the last pass is radix-16
// on *pairs* of butterflies (columns j and PP-j) so that the Hermitian split X[k] = E[k] - i w_n^k O[k]
// finds both Z[k] and Z[m-k] in the same thread's registers and the result goes to global memory
// without another exchange.  Twiddles come from shared-memory tables built once per persistent CTA.
#pragma once
#include "layout.cuh"

namespace hpxfft_b200 {

constexpr int ROW_THREADS = 256;
constexpr int ROW_PT = 32; // complex points per thread

// prefix radices (everything before the paired radix-16 pass) and shared-memory padding shift
template <int M> struct RowPlan;
template <> struct RowPlan<32>   { static constexpr int NPRE = 1, R0 = 2,  R1 = 1,  PS = 5; };
template <> struct RowPlan<64>   { static constexpr int NPRE = 1, R0 = 4,  R1 = 1,  PS = 5; };
template <> struct RowPlan<128>  { static constexpr int NPRE = 1, R0 = 8,  R1 = 1,  PS = 5; };
template <> struct RowPlan<256>  { static constexpr int NPRE = 1, R0 = 16, R1 = 1,  PS = 4; };
template <> struct RowPlan<512>  { static constexpr int NPRE = 1, R0 = 32, R1 = 1,  PS = 5; };
template <> struct RowPlan<1024> { static constexpr int NPRE = 2, R0 = 8,  R1 = 8,  PS = 3; };
template <> struct RowPlan<2048> { static constexpr int NPRE = 2, R0 = 16, R1 = 8,  PS = 4; };
template <> struct RowPlan<4096> { static constexpr int NPRE = 2, R0 = 16, R1 = 16, PS = 4; };
template <> struct RowPlan<8192> { static constexpr int NPRE = 2, R0 = 32, R1 = 16, PS = 5; };

template <int M> __host__ __device__ constexpr int row_lp() { return M + (M >> RowPlan<M>::PS); } // padded pencil
template <int M> __host__ __device__ constexpr int row_tpr() { return M / ROW_PT; }              // threads per row
template <int M> __host__ __device__ constexpr int row_group() { return ROW_THREADS / row_tpr<M>(); } // rows per CTA
template <int M> __host__ __device__ constexpr size_t row_smem_bytes() { return (size_t) row_group<M>() * row_lp<M>() * sizeof(cd); }

template <int PS> __device__ __forceinline__ int rpad(int a) { return a + (a >> PS); }

// Hermitian split of one pair: a = Z[k], b = Z[m-k], w = w_n^k  ->  xk = X[k], xmk = X[m-k]
__device__ __forceinline__ void herm_pair(cd a, cd b, cd w, cd &xk, cd &xmk)
{
#ifdef HPXFFT_B200_DIAG_NOMATH
    xk = a;
    xmk = b;
    return;
#endif
    const cd s = make_double2(a.x + b.x, a.y - b.y); // a + conj(b)
    const cd d = make_double2(a.x - b.x, a.y + b.y); // a - conj(b)
    const cd t = cmul(d, make_double2(0.5 * w.x, 0.5 * w.y));
    const cd e = make_double2(0.5 * s.x, 0.5 * s.y);
    xk = make_double2(e.x + t.y, e.y - t.x);
    xmk = make_double2(e.x - t.y, -(e.y + t.x));
}

// cp.async (LDGSTS) 16-byte global -> shared copy, L2-only caching; used to stage the NEXT row's raw
// input into the pencil buffer while the current row is still in its register-only tail
// (last-pass butterflies, Hermitian split, stores)
__device__ __forceinline__ void cp_async16(cd *dst_smem, const cd *src_gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned) __cvta_generic_to_shared(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int M> __device__ __forceinline__ void stage_row(cd *sm, const cd *__restrict__ zrow, int lt)
{
    constexpr int TPR = row_tpr<M>();
#pragma unroll
    for (int e = 0; e < ROW_PT; ++e) cp_async16(sm + lt + e * TPR, zrow + lt + e * TPR);
}

// ---- shared-memory twiddle tables of the row kernel (built once per persistent CTA) ---------------
//   tw1[r*R0 + k]  = w_{R0 R1}^(r k)          second prefix pass (k < R0, r < R1), only when NPRE == 2
//   tw2[r*JW + j]  = w_M^(r j)                last pass, r < 16, j <= PP/2 (column PP-j uses the conjugate:
//                                             w_M^(r (PP-j)) = w_16^r conj(w_M^(r j)), and the w_16^r factor
//                                             only rotates the butterfly's output index by one)
//   tw3[j]         = w_n^j                    Hermitian split, j <= PP/2; the s-dependence is w_32^s, a constant
template <int M> __host__ __device__ constexpr int row_jw() { return M / 32 + 1; }
template <int M> __host__ __device__ constexpr int row_tw1_entries() { return RowPlan<M>::NPRE == 2 ? RowPlan<M>::R0 * RowPlan<M>::R1 : 0; }
template <int M> __host__ __device__ constexpr int row_tw_entries() { return row_tw1_entries<M>() + 17 * row_jw<M>(); }
template <int M> __host__ __device__ constexpr size_t row_smem_total() { return row_smem_bytes<M>() + (size_t) row_tw_entries<M>() * sizeof(cd); }

// one prefix pass: radix R, NS = product of earlier radices (C, zrow, tw, c: unused, kept for the callers in the long-row kernels)
template <int M, int C, int R, int NS, bool FIRST>
__device__ __forceinline__ void row_pass(cd (&v)[ROW_PT], cd *sm, const cd *__restrict__ zrow, const cd *__restrict__ tw, const cd *tw1,
                                         int lt, int c)
{
    constexpr int PS = RowPlan<M>::PS;
    constexpr int NB = ROW_PT / R, T = M / R, TPR = row_tpr<M>();
    constexpr int LGR = ilog2(R);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int j = lt + b * TPR;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (FIRST)
                v[b * R + r] = sm[j + r * T]; // raw row staged by cp.async (identity layout, thread-private)
            else
                v[b * R + r] = sm[rpad<PS>(j + r * T)];
        }
    }
    __syncthreads(); // FIRST: previous row's last-pass reads are done; else: all reads before in-place writes
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int j = lt + b * TPR;
        const int k = j & (NS - 1);
        cd w[R];
#pragma unroll
        for (int r = 0; r < R; ++r) w[r] = v[b * R + r];
        if (NS > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) w[r] = cmul(w[r], tw1[r * NS + k]);
        }
        fft_dif<R>(w);
        const int j0 = ((j - k) << LGR) + k;
#pragma unroll
        for (int s = 0; s < R; ++s) sm[rpad<PS>(j0 + s * NS)] = w[bitrev(s, LGR)];
    }
    __syncthreads();
}

// tw: w_n^i, i < n = 2*M.  V rows have `pitch` complex (= cy) elements.  grid = persistent CTAs.
// Rows longer than one pencil (ny >= 32768) are handled by kernels_rows_long2.cuh / kernels_rows_long.cuh.
// FASTADDR: single destination rank and CW | PP -- output pointers advance by a constant per s.
template <int M, bool FASTADDR>
__global__ void __launch_bounds__(ROW_THREADS, 1)
    rows_r2c_kernel(const cd *__restrict__ V, unsigned pitch, unsigned nxl, RowDst dst, const cd *__restrict__ tw)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using P = RowPlan<M>;
    constexpr int PS = P::PS, TPR = row_tpr<M>(), G = row_group<M>(), LP = row_lp<M>();
    constexpr int PP = M / 16; // columns of the last pass
    constexpr int JW = row_jw<M>();
    constexpr unsigned MM = (unsigned) M; // complex length of the row
    const int g = threadIdx.x / TPR, lt = threadIdx.x % TPR;
    cd *sm = reinterpret_cast<cd *>(smem_raw) + g * LP;
    cd *tw1 = reinterpret_cast<cd *>(smem_raw + row_smem_bytes<M>());
    cd *tw2 = tw1 + row_tw1_entries<M>();
    cd *tw3 = tw2 + 16 * JW;
    const unsigned ngroups = (nxl + G - 1) / G;

    // build the twiddle tables (the first row_pass barrier publishes them)
    if constexpr (P::NPRE == 2) {
        for (int i = threadIdx.x; i < P::R0 * P::R1; i += ROW_THREADS) {
            const int r = i / P::R0, k = i - r * P::R0;
            tw1[i] = ldtw(tw, (unsigned) (r * k) * (unsigned) (2 * M / (P::R0 * P::R1)));
        }
    }
    for (int i = threadIdx.x; i < 16 * JW; i += ROW_THREADS) {
        const int r = i / JW, j = i - r * JW;
        tw2[i] = ldtw(tw, 2u * (unsigned) (r * j));
    }
    for (int i = threadIdx.x; i < JW; i += ROW_THREADS) tw3[i] = ldtw(tw, (unsigned) i);

    auto row_ptr = [&](unsigned grp) -> const cd * {
        unsigned row = grp * G + g;
#ifdef HPXFFT_B200_DIAG_WRAP
        row &= 63u;
#endif
        return V + (unsigned long long) (row < nxl ? row : nxl - 1) * pitch;
    };
    if (blockIdx.x < ngroups) stage_row<M>(sm, row_ptr(blockIdx.x), lt);
    for (unsigned grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
#ifdef HPXFFT_B200_DIAG_WRAP
        const unsigned row = (grp * G + g) & 63u;
#else
        const unsigned row = grp * G + g;
#endif
        const bool valid = row < nxl;
        const cd *zrow = row_ptr(grp);
        cd v[ROW_PT];
        cp_async_wait_all(); // own copies landed; each thread only reads what it staged itself
        row_pass<M, 1, P::R0, 1, true>(v, sm, zrow, tw, tw1, lt, 0);
        if constexpr (P::NPRE == 2) row_pass<M, 1, P::R1, P::R0, false>(v, sm, zrow, tw, tw1, lt, 0);

        // ---- last pass: two radix-16 butterflies (columns jA, jB); the partner of k2 is M - k2 ----
        const int jA = lt, jB = lt ? PP - lt : PP / 2;
        cd A[16], B[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            A[r] = sm[rpad<PS>(jA + r * PP)];
            B[r] = sm[rpad<PS>(jB + r * PP)];
        }
        // the pencil buffer is dead until the next row's first pass: refill it with the next row's
        // raw input while this row finishes in registers
        __syncthreads();
        if (grp + gridDim.x < ngroups) stage_row<M>(sm, row_ptr(grp + gridDim.x), lt);
        // column jB gets conj twiddles (w_M^(r (PP - j)) = w_16^r conj(w_M^(r j))): its natural output s sits at butterfly output (s+1)&15
        if (lt != 0) {
#pragma unroll
            for (int r = 1; r < 16; ++r) {
                const cd t = tw2[r * JW + jA];
                A[r] = cmul(A[r], t);
                B[r] = cmulc(B[r], t);
            }
        } else {
#pragma unroll
            for (int r = 1; r < 16; ++r) B[r] = mulw32(B[r], r); // jA = 0, jB = PP/2: w_M^(r PP/2) = w_32^r
        }
        fft_dif<16>(A);
        fft_dif<16>(B);
        // natural order: Z[jA + s*PP] = A[bitrev(s)],  Z[jB + s*PP] = B[bitrev(lt ? s+1 : s)]
        if (lt != 0) {
            const cd wb = tw3[jA]; // w_n^jA
            cd *pk = nullptr, *pm = nullptr;
            long long step = 0;
            if constexpr (FASTADDR) {
                const unsigned k0 = (unsigned) jA, m0 = MM - k0;
                pk = dst.base[0] + (unsigned long long) (k0 >> CW_SHIFT) * dst.tile_stride + (unsigned long long) row * CW + (k0 & (unsigned) (CW - 1));
                pm = dst.base[0] + (unsigned long long) (m0 >> CW_SHIFT) * dst.tile_stride + (unsigned long long) row * CW + (m0 & (unsigned) (CW - 1));
                step = (long long) (PP >> CW_SHIFT) * (long long) dst.tile_stride;
            }
#pragma unroll
            for (int s = 0; s < 16; ++s) {
                cd xk, xmk;
                // pair Z[kA] with Z[MM - kA] = natural output 15-s of column jB = butterfly output (16-s)&15
                herm_pair(A[bitrev(s, 4)], B[bitrev((16 - s) & 15, 4)], mulw32(wb, s), xk, xmk);
                if (valid) {
                    if constexpr (FASTADDR) {
                        st_stream(pk + s * step, xk);
                        st_stream(pm - s * step, xmk);
                    } else {
                        const unsigned kA = (unsigned) (jA + s * PP);
                        st_stream(rowdst_ptr(dst, row, kA), xk);
                        st_stream(rowdst_ptr(dst, row, MM - kA), xmk);
                    }
                }
            }
        } else {
            // lt == 0: columns 0 and PP/2 are their own partners
            const cd z0 = A[0];
            if (valid) {
                st_stream(rowdst_ptr(dst, row, 0u), make_double2(z0.x + z0.y, 0.0));
                st_stream(rowdst_ptr(dst, row, MM), make_double2(z0.x - z0.y, 0.0));
            }
#pragma unroll
            for (int s = 1; s < 8; ++s) {
                const unsigned k = (unsigned) (s * PP);
                cd xk, xmk;
                herm_pair(A[bitrev(s, 4)], A[bitrev(16 - s, 4)], mulw32(make_double2(1.0, 0.0), s), xk, xmk);
                if (valid) {
                    st_stream(rowdst_ptr(dst, row, k), xk);
                    st_stream(rowdst_ptr(dst, row, MM - k), xmk);
                }
            }
            if (valid) st_stream(rowdst_ptr(dst, row, (unsigned) (8 * PP)), cconj(A[bitrev(8, 4)]));
            const cd wh = tw3[PP / 2]; // w_n^(PP/2)
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const unsigned k = (unsigned) (PP / 2 + s * PP);
                cd xk, xmk;
                herm_pair(B[bitrev(s, 4)], B[bitrev(15 - s, 4)], mulw32(wh, s), xk, xmk);
                if (valid) {
                    st_stream(rowdst_ptr(dst, row, k), xk);
                    st_stream(rowdst_ptr(dst, row, MM - k), xmk);
                }
            }
        }
    }
}

// ---- tiny rows, m = ny/2 in {1,2,4,8,16}: one thread per row, everything in registers ----------
template <int M>
__global__ void rows_r2c_tiny_kernel(const cd *__restrict__ V, unsigned pitch, unsigned nxl, RowDst dst, const cd *__restrict__ tw)
{
    const unsigned row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nxl) return;
    const cd *zrow = V + (unsigned long long) row * pitch;
    constexpr int L = ilog2(M);
    cd v[M];
#pragma unroll
    for (int r = 0; r < M; ++r) v[r] = zrow[r];
    fft_dif<M>(v);
    const cd z0 = v[0];
    *rowdst_ptr(dst, row, 0u) = make_double2(z0.x + z0.y, 0.0);
    *rowdst_ptr(dst, row, (unsigned) M) = make_double2(z0.x - z0.y, 0.0);
#pragma unroll
    for (int s = 1; s < M / 2; ++s) {
        cd xk, xmk;
        herm_pair(v[bitrev(s, L)], v[bitrev(M - s, L)], ldtw(tw, (unsigned) s), xk, xmk);
        *rowdst_ptr(dst, row, (unsigned) s) = xk;
        *rowdst_ptr(dst, row, (unsigned) (M - s)) = xmk;
    }
    if (M >= 2) *rowdst_ptr(dst, row, (unsigned) (M / 2)) = cconj(v[bitrev(M / 2, L)]);
}

}  // namespace hpxfft_b200

// kernels_rows.cuh -- in-place-semantics 1-D r2c FFT of every local row (length ny, ny even).
//
// Replaces fft_1d_r2c_inplace -> fftw_execute_dft_r2c (core/src/shared/loop.cpp:6-9,
// core/src/distributed/loop.cpp:7-10, core/src/util/adapter_fftw.cpp:12-15) fused with the pack /
// transpose that follows it in the reference (shared/loop.cpp:18-25, distributed/loop.cpp:19-27):
// the result is written directly in the column-tiled intermediate layout (layout.cuh).
//
// Algorithm: z[j] = x[2j] + i x[2j+1] (a row of ny doubles *is* ny/2 double2), length-m complex
// Stockham FFT (m = ny/2) with the whole pencil resident in shared memory, first pass fed straight
// from coalesced 16-byte global loads, last pass radix-16 on *pairs* of butterflies (columns j and
// PP-j) so that the Hermitian split X[k] = E[k] - i w_n^k O[k] finds both Z[k] and Z[m-k] in the
// same thread's registers and the result goes to global memory without another exchange.
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
    const cd s = make_double2(a.x + b.x, a.y - b.y); // a + conj(b)
    const cd d = make_double2(a.x - b.x, a.y + b.y); // a - conj(b)
    const cd t = cmul(d, make_double2(0.5 * w.x, 0.5 * w.y));
    const cd e = make_double2(0.5 * s.x, 0.5 * s.y);
    xk = make_double2(e.x + t.y, e.y - t.x);
    xmk = make_double2(e.x - t.y, -(e.y + t.x));
}

// Long rows (m = M*C > 8192 complex): decimation in frequency over the leading index.  CTA c of a row
// computes y_c[j] = w_m^(j c) * sum_{j1<C} z[j + j1 M] w_C^(j1 c), whose M-point FFT is Z[c + C k2].
// Every CTA reads the whole row (served by L2 for all but the first reader) and no CTA talks to another.
template <int M, int C>
__device__ __forceinline__ cd load_split(const cd *__restrict__ zrow, int idx, int c, const cd *__restrict__ tw)
{
    if constexpr (C == 1) {
        return ld_stream(zrow + idx);
    } else if constexpr (C == 2) {
        const cd a = ld_stream(zrow + idx), b = ld_stream(zrow + idx + M);
        if (c == 0) return cadd(a, b);
        return cmul(csub(a, b), ldtw(tw, 2u * (unsigned) idx));
    } else {
        cd acc = ld_stream(zrow + idx);
#pragma unroll
        for (int j1 = 1; j1 < C; ++j1)
            acc = cadd(acc, cmul(ld_stream(zrow + idx + j1 * M), ldtw(tw, (unsigned) ((j1 * c) % C) * (unsigned) (2 * M))));
        return c == 0 ? acc : cmul(acc, ldtw(tw, 2u * (unsigned) idx * (unsigned) c));
    }
}

// one prefix pass: radix R, NS = product of earlier radices
template <int M, int C, int R, int NS, bool FIRST>
__device__ __forceinline__ void row_pass(cd (&v)[ROW_PT], cd *sm, const cd *__restrict__ zrow, const cd *__restrict__ tw, int lt,
                                         int c)
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
                v[b * R + r] = load_split<M, C>(zrow, j + r * T, c, tw);
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
            for (int r = 1; r < R; ++r) w[r] = cmul(w[r], ldtw(tw, (unsigned) (r * k) * (unsigned) (2 * C * M / (NS * R))));
        }
        fft_dif<R>(w);
        const int j0 = ((j - k) << LGR) + k;
#pragma unroll
        for (int s = 0; s < R; ++s) sm[rpad<PS>(j0 + s * NS)] = w[bitrev(s, LGR)];
    }
    __syncthreads();
}

// tw: w_n^i, i < n = 2*M*C.  V rows have `pitch` complex (= cy) elements.  grid = (row groups, C).
// C <= 2: Hermitian split fused (the partner of Z[c + C k2] lives in the same CTA); output via RowDst.
// C  > 2: the partner lives in CTA C-c, so the raw Z is written to `zraw` (row-major, pitch m = M*C) and
//         herm_split_kernel finishes the job.
template <int M, int C>
__global__ void __launch_bounds__(ROW_THREADS, 1)
    rows_r2c_kernel(const cd *__restrict__ V, unsigned pitch, unsigned nxl, RowDst dst, const cd *__restrict__ tw, cd *__restrict__ zraw)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using P = RowPlan<M>;
    constexpr int PS = P::PS, TPR = row_tpr<M>(), G = row_group<M>(), LP = row_lp<M>();
    constexpr int PP = M / 16; // columns of the last pass
    constexpr unsigned MM = (unsigned) M * C; // complex length of the whole row
    const int g = threadIdx.x / TPR, lt = threadIdx.x % TPR;
    const int c = C == 1 ? 0 : (int) blockIdx.y;
    cd *sm = reinterpret_cast<cd *>(smem_raw) + g * LP;
    const unsigned ngroups = (nxl + G - 1) / G;

    for (unsigned grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const unsigned row = grp * G + g;
        const bool valid = row < nxl;
        const cd *zrow = V + (unsigned long long) (valid ? row : nxl - 1) * pitch;
        cd v[ROW_PT];
        row_pass<M, C, P::R0, 1, true>(v, sm, zrow, tw, lt, c);
        if constexpr (P::NPRE == 2) row_pass<M, C, P::R1, P::R0, false>(v, sm, zrow, tw, lt, c);

        // ---- last pass: two radix-16 butterflies (columns jA, jB) ----
        // partner of k2 is M - k2 (c == 0) or M - 1 - k2 (C == 2, c == 1)
        const bool odd = (C == 2 && c == 1);
        const int jA = lt, jB = odd ? PP - 1 - lt : (lt ? PP - lt : PP / 2);
        cd A[16], B[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            A[r] = sm[rpad<PS>(jA + r * PP)];
            B[r] = sm[rpad<PS>(jB + r * PP)];
        }
#pragma unroll
        for (int r = 1; r < 16; ++r) {
            A[r] = cmul(A[r], ldtw(tw, (unsigned) (2 * C) * (unsigned) (r * jA)));
            B[r] = cmul(B[r], ldtw(tw, (unsigned) (2 * C) * (unsigned) (r * jB)));
        }
        fft_dif<16>(A);
        fft_dif<16>(B);
        // natural order: Z[c + C*(j + s*PP)] = X_[bitrev(s)]
        if constexpr (C > 2) {
            if (valid) {
                cd *zr = zraw + (unsigned long long) row * MM;
#pragma unroll
                for (int s = 0; s < 16; ++s) {
                    st_stream(zr + (unsigned) c + (unsigned) C * (unsigned) (jA + s * PP), A[bitrev(s, 4)]);
                    st_stream(zr + (unsigned) c + (unsigned) C * (unsigned) (jB + s * PP), B[bitrev(s, 4)]);
                }
            }
        } else if (odd || lt != 0) {
#pragma unroll
            for (int s = 0; s < 16; ++s) {
                const unsigned kA = (unsigned) c + (unsigned) C * (unsigned) (jA + s * PP);
                cd xk, xmk;
                herm_pair(A[bitrev(s, 4)], B[bitrev(15 - s, 4)], ldtw(tw, kA), xk, xmk);
                if (valid) {
                    st_stream(rowdst_ptr(dst, row, kA), xk);
                    st_stream(rowdst_ptr(dst, row, MM - kA), xmk);
                }
            }
        } else {
            const cd z0 = A[0];
            if (valid) {
                st_stream(rowdst_ptr(dst, row, 0u), make_double2(z0.x + z0.y, 0.0));
                st_stream(rowdst_ptr(dst, row, MM), make_double2(z0.x - z0.y, 0.0));
            }
#pragma unroll
            for (int s = 1; s < 8; ++s) {
                const unsigned k = (unsigned) C * (unsigned) (s * PP);
                cd xk, xmk;
                herm_pair(A[bitrev(s, 4)], A[bitrev(16 - s, 4)], ldtw(tw, k), xk, xmk);
                if (valid) {
                    st_stream(rowdst_ptr(dst, row, k), xk);
                    st_stream(rowdst_ptr(dst, row, MM - k), xmk);
                }
            }
            if (valid) st_stream(rowdst_ptr(dst, row, (unsigned) C * (unsigned) (8 * PP)), cconj(A[bitrev(8, 4)]));
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const unsigned k = (unsigned) C * (unsigned) (PP / 2 + s * PP);
                cd xk, xmk;
                herm_pair(B[bitrev(s, 4)], B[bitrev(15 - s, 4)], ldtw(tw, k), xk, xmk);
                if (valid) {
                    st_stream(rowdst_ptr(dst, row, k), xk);
                    st_stream(rowdst_ptr(dst, row, MM - k), xmk);
                }
            }
        }
    }
}

// Hermitian split as a separate pass (rows longer than 2*8192 complex): zraw[row][k], k < m  ->  X[k], k <= m.
// grid = (ceil((m/2+1)/256), nxl)
__global__ void herm_split_kernel(const cd *__restrict__ zraw, unsigned m, unsigned nxl, RowDst dst, const cd *__restrict__ tw)
{
    const unsigned row = blockIdx.y;
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nxl || k > m / 2) return;
    const cd *zr = zraw + (unsigned long long) row * m;
    if (k == 0) {
        const cd z0 = zr[0];
        *rowdst_ptr(dst, row, 0u) = make_double2(z0.x + z0.y, 0.0);
        *rowdst_ptr(dst, row, m) = make_double2(z0.x - z0.y, 0.0);
    } else if (2 * k == m) {
        *rowdst_ptr(dst, row, k) = cconj(zr[k]);
    } else {
        cd xk, xmk;
        herm_pair(zr[k], zr[m - k], ldtw(tw, k), xk, xmk);
        *rowdst_ptr(dst, row, k) = xk;
        *rowdst_ptr(dst, row, m - k) = xmk;
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

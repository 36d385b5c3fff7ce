// kernels_rows_long2.cuh -- r2c FFT of rows with ny = 32768 (m = 16384 complex): the row length of BASELINE configs 3 and 4.
//
// Same role as kernels_rows.cuh (fft_1d_r2c_inplace + the pack/transpose that follows it in the reference,
// core/src/distributed/loop.cpp:7-10,19-27).  One persistent CTA owns a whole row and runs its two decimation-in-frequency
// halves one after the other through the shared-memory pencil:
//   c = 0:  y_0[j] = z[j] + z[j + M]             -> Z[2 k2]     (even bins)
//   c = 1:  y_1[j] = (z[j] - z[j + M]) w_m^j     -> Z[2 k2 + 1] (odd bins)
// The Hermitian partner of a bin has the same parity, so each half finishes its own bins in registers (paired radix-16
// last pass, as in rows_r2c_kernel).  What the two halves must NOT do is store their bins separately: bins 2 k2 and 2 k2 + 1
// share a 32-byte sector, and a half-written sector costs a DRAM read-modify-write (round-1 kernel, ncu on 32768^2:
// 34.5 GB read + 17.0 GB written for 8.6 + 8.6 GB of algorithmic traffic).  So the even half parks X[2 k2] in a per-CTA
// scratch (128 KB, rewritten for every row, L2-resident, evict_last) and the odd half, which produces X[2 k2 + 1] for the
// same k2 in the same order, reads it back and emits both with one 256-bit store per lane: a warp instruction covers
// 1 KB of consecutive bins.  The row itself is read from HBM once (evict_last) and from L2 the second time (evict_first).
#pragma once
#include "kernels_rows.cuh"

namespace hpxfft_b200 {

namespace rl2 {
constexpr int M = 8192, PP = M / 16, JW = PP / 2 + 1;
constexpr int LP = row_lp<M>();
// pencil | tw1 (32*16) | tw2 [16][JW+1] w_M^(r j), j <= PP/2 + 1 | tw3e [JW] w_n^(2 j) | tw3o [JW+1] w_n^(1 + 2 j) | twc [64] w_64^i
constexpr int TW_ENTRIES = 512 + 16 * (JW + 1) + JW + (JW + 1) + 64;
constexpr size_t SMEM = (size_t) (LP + TW_ENTRIES) * sizeof(cd);
constexpr size_t SCRATCH_ELEMS = M + 8; // X[2 k2], k2 = 0..M (the Nyquist bin included)
}  // namespace rl2

// PAIRED: every destination rank boundary is even, so bins 2 k2 and 2 k2 + 1 are adjacent in the destination
template <bool FASTADDR>
__global__ void __launch_bounds__(ROW_THREADS, 1)
    rows_long2_kernel(const cd *__restrict__ V, unsigned pitch, unsigned nxl, RowDst dst, const cd *__restrict__ tw, cd *__restrict__ scratch)
{
    using namespace rl2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int PS = RowPlan<M>::PS;
    constexpr unsigned MM = 2u * M;
    const int lt = threadIdx.x;
    cd *sm = reinterpret_cast<cd *>(smem_raw);
    cd *tw1 = sm + LP;
    cd *tw2 = tw1 + 512;
    cd *tw3e = tw2 + 16 * (JW + 1);
    cd *tw3o = tw3e + JW;
    cd *twc = tw3o + (JW + 1);
    cd *xe = scratch + (size_t) blockIdx.x * SCRATCH_ELEMS;
    const unsigned long long keep = l2_policy_evict_last(), drop = l2_policy_evict_first();
    const bool pair_ok = FASTADDR || (dst.wq0 % 2u == 0u);

    // tables from tw = w_n^i, n = 2 m = 4 M
    for (int i = lt; i < 512; i += ROW_THREADS) {
        const int r = i >> 5, k = i & 31;
        tw1[i] = ldtw(tw, (unsigned) (r * k) * (unsigned) (4 * M / 512));
    }
    for (int i = lt; i < 16 * (JW + 1); i += ROW_THREADS) {
        const int r = i / (JW + 1), j = i - r * (JW + 1);
        tw2[i] = ldtw(tw, 4u * (unsigned) (r * j));
    }
    for (int i = lt; i < JW; i += ROW_THREADS) tw3e[i] = ldtw(tw, 2u * (unsigned) i);
    for (int i = lt; i < JW + 1; i += ROW_THREADS) tw3o[i] = ldtw(tw, 1u + 2u * (unsigned) i);
    for (int i = lt; i < 64; i += ROW_THREADS) twc[i] = ldtw(tw, (unsigned) i * (unsigned) (4 * M / 64));
    __syncthreads();

    auto out_ptr = [&](unsigned row, unsigned k) -> cd * {
        if constexpr (FASTADDR)
            return dst.base[0] + (unsigned long long) (k >> CW_SHIFT) * dst.tile_stride + (unsigned long long) row * CW + (k & (unsigned) (CW - 1));
        else
            return rowdst_ptr(dst, row, k);
    };
    // bins 2 k2 (parked value e) and 2 k2 + 1 (fresh value o)
    auto emit_pair = [&](unsigned row, unsigned k2, cd e, cd o) {
        if (pair_ok)
            st_stream_pair(out_ptr(row, 2u * k2), e, o);
        else {
            st_stream(out_ptr(row, 2u * k2), e);
            st_stream(out_ptr(row, 2u * k2 + 1u), o);
        }
    };

    for (unsigned row0 = blockIdx.x; row0 < nxl; row0 += gridDim.x) {
#ifdef HPXFFT_B200_DIAG_WRAP
        const unsigned row = row0 & 63u;
#else
        const unsigned row = row0;
#endif
        const cd *zrow = V + (unsigned long long) row * pitch;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            cd v[ROW_PT];
            // ---- pass 0: DIF combine of the two halves of the row, twiddle w_m^(idx c), idx = lt + 256 r, radix 32 ----
            {
                const cd wbase = c ? ldtw(tw, 2u * (unsigned) lt) : make_double2(1.0, 0.0); // w_m^lt
                const unsigned long long pol = c ? drop : keep;
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    const cd *p = zrow + lt + r * 256;
                    const cd a = ld_cg_hint(p, pol), b = ld_cg_hint(p + M, pol);
                    v[r] = c ? cmul(csub(a, b), cmul(wbase, twc[r])) : cadd(a, b); // w_m^(256 r) = w_64^r
                }
            }
            __syncthreads(); // the previous half's last-pass reads of the pencil are done
            fft_dif<32>(v);
#pragma unroll
            for (int s = 0; s < 32; ++s) sm[rpad<PS>(lt * 32 + s)] = v[bitrev(s, 5)];
            __syncthreads();
            // ---- pass 1: radix 16 ----
            row_pass<M, 1, 16, 32, false>(v, sm, zrow, tw, tw1, lt, 0);

            // ---- last pass: two radix-16 butterflies on the paired columns jA, jB; partner of k2 is M - k2 (c = 0) or M - 1 - k2 (c = 1)
            const int jA = lt, jB = c ? PP - 1 - lt : (lt ? PP - lt : PP / 2);
            cd A[16], B[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                A[r] = sm[rpad<PS>(jA + r * PP)];
                B[r] = sm[rpad<PS>(jB + r * PP)];
            }
            // odd half: the parked even bins of the A-side outputs (k2 = jA + s PP) are fetched now; the latency hides behind the butterflies
            cd EA[16];
            if (c) {
#pragma unroll
                for (int s = 0; s < 16; ++s) EA[s] = ld_cg_hint(xe + jA + s * PP, keep);
            }
            if (c) {
#pragma unroll
                for (int r = 1; r < 16; ++r) {
                    A[r] = cmul(A[r], tw2[r * (JW + 1) + jA]);
                    B[r] = cmulc(B[r], tw2[r * (JW + 1) + jA + 1]); // jB = PP - (jA + 1)
                }
            } else if (lt != 0) {
#pragma unroll
                for (int r = 1; r < 16; ++r) {
                    const cd t = tw2[r * (JW + 1) + jA];
                    A[r] = cmul(A[r], t);
                    B[r] = cmulc(B[r], t);
                }
            } else {
#pragma unroll
                for (int r = 1; r < 16; ++r) B[r] = mulw32(B[r], r); // jA = 0, jB = PP/2: w_M^(r PP/2) = w_32^r
            }
            fft_dif<16>(A);
            fft_dif<16>(B);
            // natural order: Z[c + 2 (jA + s PP)] = A[bitrev(s)];  column jB got conj twiddles: its natural output s sits at butterfly output (s+1)&15
            if (c == 0) {
                if (lt != 0) {
                    const cd wb = tw3e[jA]; // w_n^(2 jA)
#pragma unroll
                    for (int s = 0; s < 16; ++s) {
                        cd xk, xmk;
                        herm_pair(A[bitrev(s, 4)], B[bitrev((16 - s) & 15, 4)], mulw32(wb, s), xk, xmk);
                        const unsigned k2 = (unsigned) (jA + s * PP);
                        st_cg_hint(xe + k2, xk, keep);
                        st_cg_hint(xe + (M - k2), xmk, keep);
                    }
                } else {
                    const cd z0 = A[0];
                    st_cg_hint(xe + 0, make_double2(z0.x + z0.y, 0.0), keep);
                    st_cg_hint(xe + M, make_double2(z0.x - z0.y, 0.0), keep); // Nyquist bin X[m]
#pragma unroll
                    for (int s = 1; s < 8; ++s) {
                        cd xk, xmk;
                        herm_pair(A[bitrev(s, 4)], A[bitrev(16 - s, 4)], mulw32(make_double2(1.0, 0.0), s), xk, xmk);
                        st_cg_hint(xe + s * PP, xk, keep);
                        st_cg_hint(xe + (M - s * PP), xmk, keep);
                    }
                    st_cg_hint(xe + 8 * PP, cconj(A[bitrev(8, 4)]), keep);
                    const cd wh = tw3e[PP / 2]; // w_n^(2 PP/2)
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        cd xk, xmk;
                        herm_pair(B[bitrev(s, 4)], B[bitrev(15 - s, 4)], mulw32(wh, s), xk, xmk);
                        st_cg_hint(xe + (PP / 2 + s * PP), xk, keep);
                        st_cg_hint(xe + (M - (PP / 2 + s * PP)), xmk, keep);
                    }
                }
            } else {
                const cd wb = tw3o[jA]; // w_n^(1 + 2 jA)
                cd XM[16];
#pragma unroll
                for (int s = 0; s < 16; ++s) {
                    cd xk;
                    herm_pair(A[bitrev(s, 4)], B[bitrev((16 - s) & 15, 4)], mulw32(wb, s), xk, XM[s]);
                    emit_pair(row, (unsigned) (jA + s * PP), EA[s], xk);
                }
                // mirrored side: odd bin m - (1 + 2 k2) = 1 + 2 (M - 1 - k2)
#pragma unroll
                for (int s = 0; s < 16; ++s) EA[s] = ld_cg_hint(xe + (M - 1 - (jA + s * PP)), keep);
#pragma unroll
                for (int s = 0; s < 16; ++s) emit_pair(row, (unsigned) (M - 1 - (jA + s * PP)), EA[s], XM[s]);
                if (lt == 0) st_stream(out_ptr(row, MM), ld_cg_hint(xe + M, keep)); // Nyquist bin
            }
        }
        // pull the next row into L2 while the CTA drains (issued here it is free; at the start of the row it queues behind the
        // row's own loads and costs 16 %)
        if (row0 + gridDim.x < nxl) {
            const cd *nxt = V + (unsigned long long) (row0 + gridDim.x) * pitch;
#pragma unroll
            for (int i = 0; i < (int) (MM / 8) / ROW_THREADS; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (i * ROW_THREADS + lt) * 8));
        }
    }
}

}  // namespace hpxfft_b200

// kernels_rows_dit2.cuh -- r2c FFT of rows with ny = 32768 (m = 16384 complex), decimation in time.
//
// Same role and same output contract as rows_long2_kernel (kernels_rows_long2.cuh); the row length of BASELINE configs 3 and 4.
//
// m = 2 M, M = 8192.  The row splits by SAMPLE parity, z_e[j] = z[2 j], z_o[j] = z[2 j + 1], and
//     Z[k] = Ze[k] + w_m^k Zo[k],   Z[k + M] = Ze[k] - w_m^k Zo[k],   Ze = FFT_M(z_e), Zo = FFT_M(z_o).
// Each half is exactly the problem rows_r2c_v2_kernel solves (kernels_rows_v2.cuh): staged with cp.async -- here a stride-2
// gather -- while the previous half finishes in registers, transformed by warp-local in-place passes, three CTA barriers.
// What the decimation-in-frequency kernel paid for and this one does not:
//   * its first pass loaded 64 values per thread straight from global memory with nothing to overlap them with;
//   * bins 2 k2 and 2 k2 + 1 came from different halves and share a 32-byte sector, so finished even bins were parked and every
//     store was a two-bin pair.  Here a thread ends the second half holding Zo[k] and Zo[M - k] for the same k (the paired
//     columns of the last pass) and combines them with Ze[k], Ze[M - k] into X[k], X[M - k], X[k + M], X[2 M - k]: four
//     families of bins, each consecutive across the lanes of a warp, so plain 16-byte stores cover whole sectors.
// Ze is parked between the halves: 32 values per thread, written and read back by the SAME thread at [slot][thread] (coalesced,
// no index algebra, no barrier), 128 KB per CTA, rewritten for every row, L2-resident (evict_last).
#pragma once
#include "kernels_rows.cuh"

namespace hpxfft_b200 {

namespace rd2 {
constexpr int M = 8192, PP = 512, JW = PP / 2 + 1;
constexpr int LP = M;
// pencil | twA[s*32+u] = w_512^(u s) | tw2[r*JW+j] = w_M^(r j) | tw3n[j] = w_n^j | tw3m[j] = w_m^j | tw64[s] = w_64^s
constexpr int TW_ENTRIES = 512 + 16 * JW + 2 * JW + 16;
constexpr size_t SMEM = (size_t) (LP + TW_ENTRIES) * sizeof(cd);
constexpr size_t SCRATCH_ELEMS = 32 * ROW_THREADS;
__device__ __forceinline__ int pad(int p) { return p ^ (((p >> 4) ^ (p >> 9)) & 7); } // as rv2::pad
__device__ __forceinline__ void cp_async16_hint(cd *dst_smem, const cd *src_gmem, unsigned long long policy)
{
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"((unsigned) __cvta_generic_to_shared(dst_smem)), "l"(src_gmem),
                 "l"(policy)
                 : "memory");
}
// The 8192-point core shared by the decimation-in-time kernels (same algebra as rows_r2c_v2_kernel, kernels_rows_v2.cuh).
// passes_ab: the two warp-local in-place passes over the 16 stride-16 sub-sequences of the staged pencil.
__device__ __forceinline__ void passes_ab(cd *sm, const cd *twA, int warp, int lane)
{
    // ---- pass A: radix 16 over j2 = u + 32 r for the two sub-sequences of this warp, in place ----
    {
        const int u = lane;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int j1 = 2 * warp + g;
            cd *base = sm + ((j1 & 8) + 16 * u); // element j1 + 16 u + 512 r sits at (that & ~7) | ((j1 ^ u ^ r) & 7)
            const int x = (j1 ^ u) & 7;
            cd a[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) a[r] = base[512 * r + (x ^ (r & 7))];
            fft_dif<16>(a);
#pragma unroll
            for (int s = 0; s < 16; ++s) {
                cd o = a[bitrev(s, 4)];
                if (s) o = cmul(o, twA[s * 32 + u]);
                base[512 * s + (x ^ (s & 7))] = o;
            }
        }
    }
    __syncwarp(); // pass B reads what the other lanes of this warp just wrote
    // ---- pass B: radix 32 over u for fixed s: one lane per (sub-sequence, s), in place ----
    {
        const int g = lane >> 4, s = lane & 15, j1 = 2 * warp + g;
        cd *base = sm + ((j1 & 8) + 512 * s);
        const int x = (j1 ^ s) & 7;
        cd c[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) c[u] = base[16 * u + (x ^ (u & 7))];
        fft_dif<32>(c);
#pragma unroll
        for (int t = 0; t < 32; ++t) base[16 * t + (x ^ (t & 7))] = c[bitrev(t, 5)]; // F_j1[s + 16 t]
    }
}
// last pass, part 1: the paired columns jA = lt and jB = PP - lt (lt = 0: PP/2) leave the pencil
__device__ __forceinline__ void load_columns(const cd *sm, int lt, cd (&A)[16], cd (&B)[16])
{
    const int jA = lt, jB = lt ? PP - lt : PP / 2;
    const cd *pa = sm + (16 * (jA >> 4) + 512 * (jA & 15));
    const cd *pb = sm + (16 * (jB >> 4) + 512 * (jB & 15));
    const int xa = ((jA >> 4) ^ jA) & 7, xb = ((jB >> 4) ^ jB) & 7;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        A[r] = pa[(r & 8) + (xa ^ (r & 7))];
        B[r] = pb[(r & 8) + (xb ^ (r & 7))];
    }
}
// last pass, part 2: radix 16 over j1.  Column jB gets the conjugate twiddles (w_M^(r (PP - j)) = w_16^r conj(w_M^(r j))): its
// natural output s sits at butterfly output (s + 1) & 15; column jA (and both columns of lt = 0) at butterfly output s
// (register bitrev(., 4)).
__device__ __forceinline__ void finish_columns(const cd *tw2, int lt, cd (&A)[16], cd (&B)[16])
{
    if (lt != 0) {
#pragma unroll
        for (int r = 1; r < 16; ++r) {
            const cd t = tw2[r * JW + lt];
            A[r] = cmul(A[r], t);
            B[r] = cmulc(B[r], t);
        }
    } else {
#pragma unroll
        for (int r = 1; r < 16; ++r) B[r] = mulw32(B[r], r); // jA = 0, jB = PP/2: w_M^(r PP/2) = w_32^r
    }
    fft_dif<16>(A);
    fft_dif<16>(B);
}
// X[k], X[2M-k], X[k+M], X[M-k] from Ze/Zo at k and at M-k;  wm = w_m^k, wn = w_n^k (n = 4 M)
__device__ __forceinline__ void combine4(cd ze_k, cd zo_k, cd ze_mk, cd zo_mk, cd wm, cd wn, cd &xk, cd &x2mk, cd &xkM, cd &xMk)
{
    const cd p = cmul(zo_k, wm);   // w_m^k Zo[k]
    const cd q = cmulc(zo_mk, wm); // -w_m^(M-k) Zo[M-k]   (w_m^(M-k) = -conj(w_m^k))
    herm_pair(cadd(ze_k, p), cadd(ze_mk, q), wn, xk, x2mk);                          // Z[k],   Z[2M-k]
    herm_pair(csub(ze_k, p), csub(ze_mk, q), make_double2(wn.y, -wn.x), xkM, xMk);   // Z[k+M], Z[M-k];  w_n^(k+M) = -i w_n^k
}
}  // namespace rd2

// PF: bulk L2 prefetch of the next row of this CTA (see rows_r2c_v2_kernel)
template <bool FASTADDR, bool PF = false>
__global__ void __launch_bounds__(ROW_THREADS, 1)
    rows_dit2_kernel(const cd *__restrict__ V, unsigned pitch, unsigned nxl, RowDst dst, const cd *__restrict__ tw, cd *__restrict__ scratch)
{
    using namespace rd2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr unsigned UM = (unsigned) M, MM = 2u * UM;
    const int lt = threadIdx.x, warp = lt >> 5, lane = lt & 31;
    cd *sm = reinterpret_cast<cd *>(smem_raw);
    cd *twA = sm + LP;
    cd *tw2 = twA + 512;
    cd *tw3n = tw2 + 16 * JW;
    cd *tw3m = tw3n + JW;
    cd *tw64 = tw3m + JW;
    cd *xe = scratch + (size_t) blockIdx.x * SCRATCH_ELEMS + lt; // slot i of this thread: xe[i * ROW_THREADS]
    const unsigned long long keep = l2_policy_evict_last(), drop = l2_policy_evict_first();

    // tables from tw = w_n^i, n = 2 m = 4 M (published by the first barrier of the row loop)
    for (int i = lt; i < 512; i += ROW_THREADS) {
        const int s = i >> 5, u = i & 31;
        twA[i] = ldtw(tw, (unsigned) (u * s) * (unsigned) (4 * M / 512));
    }
    for (int i = lt; i < 16 * JW; i += ROW_THREADS) {
        const int r = i / JW, j = i - r * JW;
        tw2[i] = ldtw(tw, 4u * (unsigned) (r * j));
    }
    for (int i = lt; i < JW; i += ROW_THREADS) {
        tw3n[i] = ldtw(tw, (unsigned) i);
        tw3m[i] = ldtw(tw, 2u * (unsigned) i);
    }
    if (lt < 16) tw64[lt] = ldtw(tw, (unsigned) lt * (unsigned) (4 * M / 64));

    // stride-2 gather of one parity of the row: thread lt copies elements lt + 256 e of the sub-sequence.  The first visit
    // brings the sectors in from HBM and keeps them (evict_last) for the second, which hands them back (evict_first).
    auto stage = [&](const cd *zrow, int h) {
        const unsigned long long pol = h ? drop : keep;
#pragma unroll
        for (int e = 0; e < ROW_PT; ++e) {
            const int p = lt + e * ROW_THREADS;
            cp_async16_hint(sm + pad(p), zrow + 2 * p + h, pol);
        }
    };
    auto out_ptr = [&](unsigned row, unsigned k) -> cd * {
        if constexpr (FASTADDR)
            return dst.base[0] + (unsigned long long) (k >> CW_SHIFT) * dst.tile_stride + (unsigned long long) row * CW + (k & (unsigned) (CW - 1));
        else
            return rowdst_ptr(dst, row, k);
    };
    if (blockIdx.x < nxl) stage(V + (unsigned long long) blockIdx.x * pitch, 0);

    for (unsigned row = blockIdx.x; row < nxl; row += gridDim.x) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            cp_async_wait_all();
            __syncthreads(); // (1) the sub-sequence has landed and is visible to every warp
            if constexpr (PF) {
                // the whole next row (both parities), issued while this row's second half is being transformed
                if (h == 1 && lane == 0 && row + gridDim.x < nxl)
                    l2_prefetch_bulk(V + (unsigned long long) (row + gridDim.x) * pitch + warp * (2 * M / 8), (unsigned) (2 * M / 8 * sizeof(cd)));
            }

            passes_ab(sm, twA, warp, lane);
            __syncthreads(); // (2) all 16 sub-spectra are complete

            // ---- last pass: radix 16 over j1 on the paired columns jA = lt and jB = PP - lt (lt = 0: PP/2) ----
            const int jA = lt;
            cd A[16], B[16];
            load_columns(sm, lt, A, B);
            __syncthreads(); // (3) the pencil is dead: refill it with the other parity / the next row
            if (h == 0)
                stage(V + (unsigned long long) row * pitch, 1);
            else if (row + gridDim.x < nxl)
                stage(V + (unsigned long long) (row + gridDim.x) * pitch, 0);

            finish_columns(tw2, lt, A, B);

            if (h == 0) {
                // park Ze in butterfly order; the second half reads slot i next to its own A[i] / B[i]
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    st_cg_hint(xe + i * ROW_THREADS, A[i], keep);
                    st_cg_hint(xe + (16 + i) * ROW_THREADS, B[i], keep);
                }
                continue;
            }

            if (lt != 0) {
                // k = jA + s PP pairs with M - k = jB + (15 - s) PP = butterfly output (16 - s) & 15 of column jB
                const cd wmb = tw3m[jA], wnb = tw3n[jA];
                cd *pk = nullptr, *p2 = nullptr, *pM = nullptr, *pr = nullptr;
                long long step = 0;
                if constexpr (FASTADDR) {
                    pk = out_ptr(row, (unsigned) jA);       // X[k]       ascending with s
                    p2 = out_ptr(row, MM - (unsigned) jA);  // X[2M - k]  descending
                    pM = out_ptr(row, UM + (unsigned) jA);  // X[k + M]   ascending
                    pr = out_ptr(row, UM - (unsigned) jA);  // X[M - k]   descending
                    step = (long long) (PP >> CW_SHIFT) * (long long) dst.tile_stride;
                }
                cd EA[2][4], EB[2][4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    EA[0][q] = ld_cg_hint(xe + bitrev(q, 4) * ROW_THREADS, keep);
                    EB[0][q] = ld_cg_hint(xe + (16 + bitrev((16 - q) & 15, 4)) * ROW_THREADS, keep);
                }
#pragma unroll
                for (int qt = 0; qt < 4; ++qt) {
                    if (qt < 3) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int s = 4 * (qt + 1) + q;
                            EA[(qt + 1) & 1][q] = ld_cg_hint(xe + bitrev(s, 4) * ROW_THREADS, keep);
                            EB[(qt + 1) & 1][q] = ld_cg_hint(xe + (16 + bitrev((16 - s) & 15, 4)) * ROW_THREADS, keep);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int s = 4 * qt + q;
                        cd xk, x2mk, xkM, xMk;
                        combine4(EA[qt & 1][q], A[bitrev(s, 4)], EB[qt & 1][q], B[bitrev((16 - s) & 15, 4)], mulw32(wmb, s), cmul(wnb, tw64[s]), xk,
                                 x2mk, xkM, xMk);
                        if constexpr (FASTADDR) {
                            st_stream(pk + s * step, xk);
                            st_stream(p2 - s * step, x2mk);
                            st_stream(pM + s * step, xkM);
                            st_stream(pr - s * step, xMk);
                        } else {
                            const unsigned k = (unsigned) (jA + s * PP);
                            st_stream(rowdst_ptr(dst, row, k), xk);
                            st_stream(rowdst_ptr(dst, row, MM - k), x2mk);
                            st_stream(rowdst_ptr(dst, row, UM + k), xkM);
                            st_stream(rowdst_ptr(dst, row, UM - k), xMk);
                        }
                    }
                }
            } else {
                // columns 0 and PP/2 are their own partners (non-conjugated twiddles: natural output s = butterfly output bitrev(s))
                auto ea = [&](int s) { return ld_cg_hint(xe + bitrev(s, 4) * ROW_THREADS, keep); };
                auto eb = [&](int s) { return ld_cg_hint(xe + (16 + bitrev(s, 4)) * ROW_THREADS, keep); };
                {
                    const cd ze = ea(0), zo = A[0];
                    const cd z0 = cadd(ze, zo), zM = csub(ze, zo); // Z[0], Z[M]
                    st_stream(out_ptr(row, 0u), make_double2(z0.x + z0.y, 0.0));
                    st_stream(out_ptr(row, MM), make_double2(z0.x - z0.y, 0.0)); // Nyquist bin
                    st_stream(out_ptr(row, UM), cconj(zM));                      // k = m/2
                }
#pragma unroll
                for (int s = 1; s < 8; ++s) {
                    const unsigned k = (unsigned) (s * PP);
                    cd xk, x2mk, xkM, xMk;
                    combine4(ea(s), A[bitrev(s, 4)], ea(16 - s), A[bitrev(16 - s, 4)], mulw32(make_double2(1.0, 0.0), s), tw64[s], xk, x2mk, xkM, xMk);
                    st_stream(out_ptr(row, k), xk);
                    st_stream(out_ptr(row, MM - k), x2mk);
                    st_stream(out_ptr(row, UM + k), xkM);
                    st_stream(out_ptr(row, UM - k), xMk);
                }
                {
                    // k = M/2 is its own partner: only X[M/2] and X[3M/2]
                    const cd ze = ea(8), zo = A[bitrev(8, 4)];
                    cd xk, x2mk, xkM, xMk;
                    combine4(ze, zo, ze, zo, mulw32(make_double2(1.0, 0.0), 8), tw64[8], xk, x2mk, xkM, xMk);
                    st_stream(out_ptr(row, (unsigned) (8 * PP)), xk);
                    st_stream(out_ptr(row, MM - (unsigned) (8 * PP)), x2mk);
                }
                const cd wmh = tw3m[PP / 2], wnh = tw3n[PP / 2];
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const unsigned k = (unsigned) (PP / 2 + s * PP);
                    cd xk, x2mk, xkM, xMk;
                    combine4(eb(s), B[bitrev(s, 4)], eb(15 - s), B[bitrev(15 - s, 4)], mulw32(wmh, s), cmul(wnh, tw64[s]), xk, x2mk, xkM, xMk);
                    st_stream(out_ptr(row, k), xk);
                    st_stream(out_ptr(row, MM - k), x2mk);
                    st_stream(out_ptr(row, UM + k), xkM);
                    st_stream(out_ptr(row, UM - k), xMk);
                }
            }
        }
    }
}

}  // namespace hpxfft_b200

// kernels_rows_long.cuh -- r2c FFT of rows longer than one shared-memory pencil (ny/2 = m = 8192*C, C = 2, 4, 8).
//
// Same role as kernels_rows.cuh (fft_1d_r2c_inplace + the pack/transpose that follows it in the reference,
// core/src/distributed/loop.cpp:7-10,19-27), for ny = 32768 ... 131072 -- the row length of BASELINE configs 3-5.
//
// One persistent CTA owns a whole row.  Decimation in frequency over the leading index splits the length-m complex
// FFT into C sub-FFTs of M = 8192 points, y_c[j] = w_m^(jc) sum_j1 z[j + j1 M] w_C^(j1 c), whose spectrum is
// Z[c + C k2].  The CTA runs the C sub-FFTs one after the other through the shared-memory pencil (the row itself is
// read from HBM once; the C-1 re-reads hit L2) and parks the raw sub-spectra contiguously in a per-CTA scratch
// that never leaves L2.  A final assembly loop reads Z[k] and Z[m-k] back (L2), applies the Hermitian split and
// stores X[k] and X[m-k] -- 32 consecutive bins per warp instruction, every 32-byte sector written whole.
//
// What this replaces (round 1): C CTAs per row, each re-reading the row from HBM and storing the bins k = c (mod C)
// it owned as isolated 16-byte elements.  ncu on 32768^2: 34.5 GB read + 17.0 GB written for 8.6 + 8.6 GB of
// algorithmic traffic (two HBM reads of every row, a read-modify-write fill for every half-written sector); for C > 2
// additionally a full HBM round trip through a raw-spectrum buffer and a separate Hermitian-split kernel.
#pragma once
#include "kernels_rows.cuh"

namespace hpxfft_b200 {

constexpr int LONG_M = 8192;
constexpr int LONG_LP = row_lp<LONG_M>();
// pencil | tw1 (second pass, 32*16) | tw2 (last pass, [16][256]) | twc (w_{32C}^i, i < 32C)
template <int C> __host__ __device__ constexpr size_t rows_long_smem_bytes()
{
    return (size_t) (LONG_LP + 512 + 16 * 256 + 32 * C) * sizeof(cd);
}
template <int C> __host__ __device__ constexpr size_t rows_long_scratch_elems() { return (size_t) LONG_M * C; }

template <int C>
__global__ void __launch_bounds__(ROW_THREADS, 1)
    rows_long_kernel(const cd *__restrict__ V, unsigned pitch, unsigned nxl, RowDst dst, const cd *__restrict__ tw, cd *__restrict__ scratch)
{
    static_assert(C == 2 || C == 4 || C == 8, "C = m / 8192");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int M = LONG_M, PS = RowPlan<M>::PS;
    constexpr unsigned MM = (unsigned) M * C;
    const int lt = threadIdx.x;
    cd *sm = reinterpret_cast<cd *>(smem_raw);
    cd *tw1 = sm + LONG_LP;
    cd *tw2 = tw1 + 512;
    cd *twc = tw2 + 16 * 256;
    cd *zs = scratch + (size_t) blockIdx.x * MM; // [c][k2], rewritten for every row: stays in L2
    const unsigned long long keep = l2_policy_evict_last(), drop = l2_policy_evict_first();

    // tables (tw = w_n^i, n = 2 m):  w_512^(r k) = w_n^(r k n/512),  w_M^(r j) = w_n^(2 C r j),  w_{32C}^i = w_n^(512 i)
    for (int i = lt; i < 512; i += ROW_THREADS) {
        const int r = i >> 5, k = i & 31;
        tw1[i] = ldtw(tw, (unsigned) (r * k) * (unsigned) (2 * C * M / 512));
    }
    for (int i = lt; i < 16 * 256; i += ROW_THREADS) {
        const int r = i >> 8, j = i & 255;
        tw2[i] = ldtw(tw, (unsigned) (2 * C) * (unsigned) (r * j));
    }
    for (int i = lt; i < 32 * C; i += ROW_THREADS) twc[i] = ldtw(tw, 512u * (unsigned) i);
    __syncthreads();

    for (unsigned row = blockIdx.x; row < nxl; row += gridDim.x) {
        const cd *zrow = V + (unsigned long long) row * pitch;
        for (int c = 0; c < C; ++c) {
            cd v[ROW_PT];
            // ---- pass 0: DIF combine over the C blocks of the row, twiddle, radix 32 ----
            // y_c[idx], idx = lt + 256 r:  w_m^(idx c) = w_m^(lt c) * w_{32C}^(r c)
            const cd wbase = c ? ldtw(tw, 2u * (unsigned) lt * (unsigned) c) : make_double2(1.0, 0.0);
            const unsigned long long rowpol = c == C - 1 ? drop : keep; // the row is read C times: keep it in L2 until the last one
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const cd *p = zrow + lt + r * 256;
                cd acc;
                if constexpr (C == 2) {
                    const cd a = ld_cg_hint(p, rowpol), b = ld_cg_hint(p + M, rowpol);
                    acc = c ? csub(a, b) : cadd(a, b);
                } else {
                    acc = ld_cg_hint(p, rowpol);
#pragma unroll
                    for (int j1 = 1; j1 < C; ++j1) {
                        const cd z = ld_cg_hint(p + j1 * M, rowpol);
                        acc = c ? cadd(acc, cmul(z, twc[(32 * j1 * c) & (32 * C - 1)])) : cadd(acc, z);
                    }
                }
                v[r] = c ? cmul(acc, cmul(wbase, twc[(r * c) & (32 * C - 1)])) : acc;
            }
            __syncthreads(); // the previous sub-FFT's last-pass reads of the pencil are done
            fft_dif<32>(v);
#pragma unroll
            for (int s = 0; s < 32; ++s) sm[rpad<PS>(lt * 32 + s)] = v[bitrev(s, 5)];
            __syncthreads();
            // ---- pass 1: radix 16 ----
            row_pass<M, 1, 16, 32, false>(v, sm, zrow, tw, tw1, lt, 0);
            // ---- pass 2: radix 16, raw spectrum Z[c + C k2], k2 = j + 512 s, parked in the L2 scratch ----
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int j = lt + b * 256;
                cd w[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) w[r] = sm[rpad<PS>(j + r * 512)];
#pragma unroll
                for (int r = 1; r < 16; ++r) {
                    w[r] = cmul(w[r], tw2[r * 256 + lt]);       // w_M^(r lt)
                    if (b) w[r] = mulw32(w[r], r);              // w_M^(256 r) = w_32^r
                }
                fft_dif<16>(w);
                cd *zc = zs + (size_t) c * M + j;
#pragma unroll
                for (int s = 0; s < 16; ++s) st_cg_hint(zc + s * 512, w[bitrev(s, 4)], keep);
            }
        }
        __syncthreads(); // the whole raw spectrum is in the scratch (global writes are visible block-wide after the barrier)
        if (row + gridDim.x < nxl) {
            // pull the next row into L2 while this one is assembled
            const cd *nxt = V + (unsigned long long) (row + gridDim.x) * pitch;
#pragma unroll
            for (int i = 0; i < (int) (MM / 8) / ROW_THREADS; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (i * ROW_THREADS + lt) * 8));
        }
        // ---- assembly: X[k] = E[k] - i w_n^k O[k] from Z[k], Z[m-k]; 32 consecutive bins per warp instruction ----
        {
            const unsigned warp = (unsigned) lt >> 5, lane = (unsigned) lt & 31u;
            auto zat = [&](unsigned q) -> cd {
                q = q == MM ? 0u : q;
                return ld_cg_hint(zs + (size_t) (q % C) * M + q / C, keep);
            };
            // UB blocks per batch: all loads of a batch are issued before the first Hermitian split, so that the loop
            // pays the L2 latency once per batch instead of once per block
            constexpr int UB = 8;
            static_assert((MM / 64) % (UB * (ROW_THREADS / 32)) == 0, "blocks per warp must be a multiple of the batch");
            for (unsigned blk0 = warp * UB; blk0 < MM / 64; blk0 += UB * (ROW_THREADS / 32)) {
                cd za[UB], zb[UB], wk[UB], ya, yb, wy;
#pragma unroll
                for (int u = 0; u < UB; ++u) {
                    const unsigned k = (blk0 + u) * 32u + lane;
                    za[u] = zat(k);
                    zb[u] = zat(MM - k);
                    wk[u] = ldtw(tw, k);
                }
                // lane 0 completes every mirrored block with bin m-k0-32, the mirror of k0+32; lanes u < UB fetch those
                // operands (one block each) and hand them to lane 0 through shuffles below
                if (lane < UB) {
                    const unsigned k2 = (blk0 + lane) * 32u + 32u;
                    ya = zat(k2);
                    yb = zat(MM - k2);
                    wy = ldtw(tw, k2);
                } else {
                    ya = yb = wy = make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < UB; ++u) {
                    const unsigned k0 = (blk0 + u) * 32u, k = k0 + lane;
                    cd xk, xmk;
                    herm_pair(za[u], zb[u], wk[u], xk, xmk);
                    st_stream(rowdst_ptr(dst, row, k), xk);
                    // operands of the extra bin, broadcast from lane u
                    cd ea, eb, ew;
                    ea.x = __shfl_sync(0xffffffffu, ya.x, u); ea.y = __shfl_sync(0xffffffffu, ya.y, u);
                    eb.x = __shfl_sync(0xffffffffu, yb.x, u); eb.y = __shfl_sync(0xffffffffu, yb.y, u);
                    ew.x = __shfl_sync(0xffffffffu, wy.x, u); ew.y = __shfl_sync(0xffffffffu, wy.y, u);
                    unsigned mb = MM - k;
                    if (lane == 0) {
                        if (k0 == 0) st_stream(rowdst_ptr(dst, row, MM), xmk); // Nyquist bin: the lone last column
                        cd x2;
                        herm_pair(ea, eb, ew, x2, xmk);
                        mb = MM - (k0 + 32u);
                    }
                    st_stream(rowdst_ptr(dst, row, mb), xmk);
                }
            }
        }
        __syncthreads(); // assembly reads are done before the next row overwrites the scratch
    }
}

}  // namespace hpxfft_b200

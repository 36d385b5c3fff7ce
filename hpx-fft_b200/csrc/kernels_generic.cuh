// kernels_generic.cuh -- lengths that are not powers of two: n = t * q, t odd, q = 2^a.
//
// FFTW accepts every length (core/src/util/adapter_fftw.cpp:6-10,24-30), the reference's default example is 8 x 14
// (examples/hpxfft/shared_loop_2d.cpp:142-143) and its shared weak-scaling sweep runs 512 * threads for 1..32 threads
// (benchmark/shared_benchmark.sh:100-102).  Mixed radix: the odd factor is a direct DFT (t multiply-adds per output,
// twiddles from a t-entry table), the power-of-two factor runs on the Stockham kernels / an in-place radix-4 loop.
//
//   columns  nx = t q:  x = x1 q + x2,  kx = k1 + t k2
//            cols_odd_kernel:  A[k1][x2] = w_nx^(k1 x2) sum_x1 w_t^(x1 k1) Y[x1 q + x2]        (one pass over HBM)
//            then the ordinary column kernels on t "virtual strips" per strip (length q, ColDst::vt = t) -- or nothing when q = 1
//   rows     m = ny/2 = t q <= 8192:  one CTA per row, the row in shared memory
//            sub-sequences y_j1[j2] = z[j1 + t j2]: in-place decimation-in-frequency FFT_q (radix 4, one radix-2 pass when
//            log2 q is odd), then for every k2 a radix-t combine  Z[k2 + q k1] = sum_j1 w_t^(j1 k1) w_m^(j1 k2) F_j1[k2],
//            then the Hermitian split of the r2c transform and the transposed store, as in kernels_rows.cuh.
#pragma once
#include "kernels_rows.cuh"

namespace hpxfft_b200 {

constexpr int GEN_THREADS = 256;
constexpr int GEN_TMAX_ROWS = 32;  // odd factor of a row length: inputs of one combine live in registers
constexpr int GEN_TMAX_COLS = 127; // odd factor of a column length: inputs live in shared memory

// ---- columns: odd-radix pre-stage ------------------------------------------------------------------------------------
// grid = (ceil(q / xb), ntiles); shared: tile [t][xb][CW] + w_t table [t].  direct = true (q == 1): results go straight to `out`.
struct ColsOddArgs {
    unsigned t, q, xb, nx;
    cd *S1; // [strip * t + k1][x2][c]
};

__global__ void __launch_bounds__(GEN_THREADS) cols_odd_kernel(InterView in, ColDst out, ColsOddArgs a, const cd *__restrict__ tw, bool direct)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *tile = reinterpret_cast<cd *>(smem_raw);
    cd *wt = tile + (size_t) a.t * a.xb * CW;
    const unsigned ct = blockIdx.y, x2_0 = blockIdx.x * a.xb;
    const unsigned nxb = a.q - x2_0 < a.xb ? a.q - x2_0 : a.xb;
    for (unsigned i = threadIdx.x; i < a.t; i += GEN_THREADS) wt[i] = ldtw(tw, i * a.q); // w_t^i = w_nx^(i q)
    for (unsigned i = threadIdx.x; i < a.t * nxb * CW; i += GEN_THREADS) {
        const unsigned c = i % CW, xb = (i / CW) % nxb, x1 = i / (CW * nxb);
        tile[(x1 * a.xb + xb) * CW + c] = ld_stream(inter_ptr(in, x1 * a.q + x2_0 + xb, ct, c));
    }
    __syncthreads();
    for (unsigned o = threadIdx.x; o < a.t * nxb * CW; o += GEN_THREADS) {
        const unsigned c = o % CW, xb = (o / CW) % nxb, k1 = o / (CW * nxb);
        const unsigned x2 = x2_0 + xb;
        double re = 0.0, im = 0.0;
        unsigned e = 0; // (x1 * k1) mod t
        for (unsigned x1 = 0; x1 < a.t; ++x1) {
            const cd y = tile[(x1 * a.xb + xb) * CW + c], w = wt[e];
            re += y.x * w.x - y.y * w.y;
            im += y.x * w.y + y.y * w.x;
            e += k1;
            if (e >= a.t) e -= a.t;
        }
        cd v = make_double2(re, im);
        if (direct) {
            const unsigned kl = ct * CW + c;
            if (kl < out.w) st_stream(coldst_ptr(out, k1, kl), v);
        } else {
            v = cmul(v, ldtw(tw, (unsigned) (((unsigned long long) k1 * x2) % a.nx)));
            a.S1[(((unsigned long long) ct * a.t + k1) * a.q + x2) * CW + c] = v;
        }
    }
}

// ---- rows: m = t * q in shared memory --------------------------------------------------------------------------------
// position of F[k] after the in-place DIF passes (radices 4,4,...,(2)): digit reversal in that mixed radix
__device__ __forceinline__ unsigned gen_rev(unsigned k, unsigned q, unsigned lg)
{
    // passes: lg/2 radix-4 passes, then one radix-2 pass when lg is odd.  pass p takes the p-th LOW digit of k to the
    // p-th HIGH position.
    unsigned pos = 0, len = q;
    for (unsigned p = 0; p < lg / 2; ++p) {
        len >>= 2;
        pos += (k & 3u) * len;
        k >>= 2;
    }
    if (lg & 1u) pos += (k & 1u); // len == 2 -> half == 1
    return pos;
}

struct RowsMixedArgs {
    unsigned t, q, lg; // m = t * q, q = 2^lg
};

// TT: compile-time bound of the odd factor (inputs of one combine live in TT registers); NT threads per CTA
template <int TT, int NT>
__global__ void __launch_bounds__(NT, 1)
    rows_mixed_kernel(const cd *__restrict__ V, unsigned pitch, unsigned nxl, RowDst dst, RowsMixedArgs a, const cd *__restrict__ tw)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned t = a.t, q = a.q, m = t * q;
    cd *sm = reinterpret_cast<cd *>(smem_raw);
    cd *wt = sm + m; // w_t^i
    cd *wq = wt + t; // w_q^e (twiddles of the radix-4 passes)
    const unsigned tid = threadIdx.x;
    // tw = w_n^i, n = 2 m:  w_t^i = w_n^(2 q i),  w_q^i = w_n^(2 t i),  w_m^i = w_n^(2 i)
    for (unsigned i = tid; i < t; i += NT) wt[i] = ldtw(tw, 2u * q * i);
    for (unsigned i = tid; i < q; i += NT) wq[i] = ldtw(tw, 2u * t * i);
    for (unsigned row = blockIdx.x; row < nxl; row += gridDim.x) {
        const cd *zrow = V + (unsigned long long) row * pitch;
        __syncthreads(); // previous row's split has finished reading the pencil (and wt is published)
        for (unsigned p = tid; p < m; p += NT) sm[p] = ld_stream(zrow + p);
        __syncthreads();
        // ---- FFT_q of the t sub-sequences (stride t), in place, DIF ----
        unsigned len = q;
        for (unsigned pass = 0; pass < a.lg / 2; ++pass) {
            const unsigned quarter = len >> 2, tws = q / len; // w_len^e = w_q^(e tws)
            for (unsigned idx = tid; idx < t * (q >> 2); idx += NT) {
                const unsigned j1 = idx % t, b = idx / t;
                const unsigned blk = b / quarter, i = b - blk * quarter;
                cd *p0 = sm + j1 + (size_t) t * (blk * len + i);
                const size_t st = (size_t) t * quarter;
                const cd a0 = p0[0], a1 = p0[st], a2 = p0[2 * st], a3 = p0[3 * st];
                const cd s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
                const cd md13 = make_double2(d13.y, -d13.x); // -i * d13
                cd y0 = cadd(s02, s13), y1 = cadd(d02, md13), y2 = csub(s02, s13), y3 = csub(d02, md13);
                if (i) {
                    const unsigned e = tws * i; // w_len^i = w_q^(tws i), 3 e < q
                    y1 = cmul(y1, wq[e]);
                    y2 = cmul(y2, wq[2u * e]);
                    y3 = cmul(y3, wq[3u * e]);
                }
                p0[0] = y0;
                p0[st] = y1;
                p0[2 * st] = y2;
                p0[3 * st] = y3;
            }
            __syncthreads();
            len = quarter;
        }
        if (a.lg & 1u) { // len == 2
            for (unsigned idx = tid; idx < t * (q >> 1); idx += NT) {
                const unsigned j1 = idx % t, b = idx / t;
                cd *p0 = sm + j1 + (size_t) t * (2u * b);
                const cd a0 = p0[0], a1 = p0[t];
                p0[0] = cadd(a0, a1);
                p0[t] = csub(a0, a1);
            }
            __syncthreads();
        }
        // ---- radix-t combine: Z[k2 + q k1] = sum_j1 w_t^(j1 k1) (w_m^(j1 k2) F_j1[k2]), in place over the t slots of k2 ----
        if (t > 1) {
            for (unsigned k2 = tid; k2 < q; k2 += NT) {
                cd *slot = sm + (size_t) t * gen_rev(k2, q, a.lg);
                cd in[TT];
#pragma unroll
                for (int j1 = 0; j1 < TT; ++j1) {
                    in[j1] = make_double2(0.0, 0.0);
                    if ((unsigned) j1 < t) {
                        in[j1] = slot[j1];
                        if (j1 && k2) in[j1] = cmul(in[j1], ldtw(tw, 2u * (unsigned) j1 * k2)); // j1 k2 < m
                    }
                }
                for (unsigned k1 = 0; k1 < t; ++k1) {
                    double re = 0.0, im = 0.0;
                    unsigned e = 0;
#pragma unroll
                    for (int j1 = 0; j1 < TT; ++j1) {
                        if ((unsigned) j1 < t) {
                            const cd w = wt[e];
                            re += in[j1].x * w.x - in[j1].y * w.y;
                            im += in[j1].x * w.y + in[j1].y * w.x;
                            e += k1;
                            if (e >= t) e -= t;
                        }
                    }
                    slot[k1] = make_double2(re, im);
                }
            }
            __syncthreads();
        }
        // ---- Hermitian split: X[k] = E[k] - i w_n^k O[k] from Z[k], Z[m-k]; Z[k2 + q k1] sits at k1 + t rev(k2) ----
        auto zat = [&](unsigned k) -> cd {
            if (k == m) k = 0;
            const unsigned k1 = k / q, k2 = k - k1 * q;
            return sm[k1 + (size_t) t * gen_rev(k2, q, a.lg)];
        };
        for (unsigned k = tid; k <= m / 2; k += NT) {
            if (k == 0) {
                const cd z0 = zat(0);
                *rowdst_ptr(dst, row, 0u) = make_double2(z0.x + z0.y, 0.0);
                *rowdst_ptr(dst, row, m) = make_double2(z0.x - z0.y, 0.0);
            } else if (2 * k == m) {
                *rowdst_ptr(dst, row, k) = cconj(zat(k));
            } else {
                cd xk, xmk;
                herm_pair(zat(k), zat(m - k), ldtw(tw, k), xk, xmk);
                st_stream(rowdst_ptr(dst, row, k), xk);
                st_stream(rowdst_ptr(dst, row, m - k), xmk);
            }
        }
    }
}

}  // namespace hpxfft_b200

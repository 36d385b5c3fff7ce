// kernels_rows_ditc.cuh -- r2c FFT of rows of ny = 2 C * 8192 real points, C = 2, 4, 8 (ny = 32768, 65536, 131072: the row
// lengths of BASELINE configs 3-5), decimation in time over C sample classes.
//
// Same role and output contract as rows_long_kernel<C> (kernels_rows_long.cuh).  m = C M complex points, M = 8192:
//     z_c[j] = z[C j + c],   Zc = FFT_M(z_c),   Z[k + M q] = sum_c w_C^(c q) w_m^(c k) Zc[k],     c, q = 0..C-1.
// One persistent CTA owns a row.  The classes run one after the other through the 8192-point core of kernels_rows_dit2.cuh
// (stride-C cp.async gather overlapped with the previous class's register tail, warp-local in-place passes, three barriers)
// and every thread parks the 32 values it ends a class with -- Zc[k] on its two paired columns -- at [class][slot][thread]
// in an L2 scratch: written and read back by the SAME thread, coalesced, no barrier.  After the last class (the gather of
// the next row's first class already in flight) the thread combines, for each of its 16 pairs (k, M - k),
//     u_c = w_m^(c k) Zc[k],  v_c = conj(w_m^(c k)) Zc[M - k],  U = DFT_C(u),  V = DFT_C(v)
//     X[k + M q], X[(M - k) + M (C - 1 - q)] = hermitian_split(U[q], V[(C - q) mod C], w_n^(k + M q))
// (tools/model_kernels.py: rows_ditc_model): 2 C families of bins, each consecutive across the lanes of a warp, so plain
// 16-byte stores cover whole sectors.
//
// What rows_long_kernel<C> paid for and this one does not: C direct passes over the whole row with 32 C exposed loads per thread
// each (no overlap), a raw spectrum parked with stride-C addresses and an assembly loop of dependent L2 reads.
#pragma once
#include "kernels_rows_dit2.cuh"

namespace hpxfft_b200 {

namespace rdc {
using rd2::JW;
using rd2::M;
using rd2::PP;
// pencil | twA[512] | tw2[16 JW] | tw3n[JW] = w_n^j | W[32 C] = w_{32C}^i
template <int C> __host__ __device__ constexpr size_t smem_bytes() { return (size_t) (M + 512 + 16 * JW + JW + 32 * C) * sizeof(cd); }
template <int C> __host__ __device__ constexpr size_t scratch_elems() { return (size_t) C * 32 * ROW_THREADS; } // = m
}  // namespace rdc

template <int C, bool FASTADDR>
__global__ void __launch_bounds__(ROW_THREADS, 1)
    rows_ditc_kernel(const cd *__restrict__ V, unsigned pitch, unsigned nxl, RowDst dst, const cd *__restrict__ tw, cd *__restrict__ scratch)
{
    using namespace rdc;
    static_assert(C == 2 || C == 4 || C == 8, "C = m / 8192");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LC = ilog2(C);
    constexpr unsigned UM = (unsigned) M;
    const int lt = threadIdx.x, warp = lt >> 5, lane = lt & 31;
    cd *sm = reinterpret_cast<cd *>(smem_raw);
    cd *twA = sm + M;
    cd *tw2 = twA + 512;
    cd *tw3n = tw2 + 16 * JW;
    cd *W = tw3n + JW;
    cd *xe = scratch + (size_t) blockIdx.x * scratch_elems<C>() + lt; // slot i of class c: xe[(c * 32 + i) * ROW_THREADS]
    const unsigned long long keep = l2_policy_evict_last(), drop = l2_policy_evict_first();

    // tables from tw = w_n^i, n = 2 m = 2 C M (published by the first barrier of the row loop)
    for (int i = lt; i < 512; i += ROW_THREADS) {
        const int s = i >> 5, u = i & 31;
        twA[i] = ldtw(tw, (unsigned) (u * s) * (unsigned) (2 * C * M / 512));
    }
    for (int i = lt; i < 16 * JW; i += ROW_THREADS) {
        const int r = i / JW, j = i - r * JW;
        tw2[i] = ldtw(tw, (unsigned) (2 * C) * (unsigned) (r * j));
    }
    for (int i = lt; i < JW; i += ROW_THREADS) tw3n[i] = ldtw(tw, (unsigned) i);
    for (int i = lt; i < 32 * C; i += ROW_THREADS) W[i] = ldtw(tw, (unsigned) i * (unsigned) (M / 16)); // n / (32 C) = M / 16

    // stride-C gather of class c: thread lt copies elements lt + 256 e of the sub-sequence.  Neighbouring classes share
    // sectors: keep them in L2 (evict_last) until the last class has been through.
    // The hint is dropped for <8, false>: ptxas 12.9 emits its hinted LDGSTS with a descriptor register that is never written
    // ("illegal instruction" on the device, seen on a B200); build.py scans the SASS of every kernel for that form.
    constexpr bool HINTED = !(C == 8 && !FASTADDR);
    auto stage = [&](const cd *zrow, int c, unsigned long long pol) {
#pragma unroll
        for (int e = 0; e < ROW_PT; ++e) {
            const int p = lt + e * ROW_THREADS;
            if constexpr (HINTED)
                rd2::cp_async16_hint(sm + rd2::pad(p), zrow + C * p + c, pol);
            else
                cp_async16(sm + rd2::pad(p), zrow + C * p + c);
        }
    };
    auto out_ptr = [&](unsigned row, unsigned k) -> cd * {
        if constexpr (FASTADDR)
            return dst.base[0] + (unsigned long long) (k >> CW_SHIFT) * dst.tile_stride + (unsigned long long) row * CW + (k & (unsigned) (CW - 1));
        else
            return rowdst_ptr(dst, row, k);
    };
    // parked values of one pair: slot_k holds Zc[k], slot_mk holds Zc[M - k]
    auto load_pair = [&](cd (&u)[C], cd (&v)[C], int slot_k, int slot_mk) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            u[c] = ld_cg(xe + (c * 32 + slot_k) * ROW_THREADS);
            v[c] = ld_cg(xe + (c * 32 + slot_mk) * ROW_THREADS);
        }
    };
    // k = j + s PP.  T[c] = w_m^(c j), wnj = w_n^j.  Emits q = 0..nq-1; last_single: the last q is its own partner (one store)
    auto emit = [&](unsigned row, cd (&u)[C], cd (&v)[C], unsigned j, int s, const cd (&T)[C], cd wnj, int nq, bool last_single) {
#pragma unroll
        for (int c = 1; c < C; ++c) {
            const cd w = cmul(T[c], W[(2 * c * s) & (32 * C - 1)]); // w_m^(c k) = w_m^(c j) w_{32C}^(2 c s)
            u[c] = cmul(u[c], w);
            v[c] = cmulc(v[c], w);
        }
        fft_dif<C>(u);
        fft_dif<C>(v);
        const cd wns = cmul(wnj, W[s]); // w_n^k
        const unsigned k = j + (unsigned) s * (unsigned) PP;
#pragma unroll
        for (int q = 0; q < C; ++q) {
            if (q < nq) {
                cd xk, xmk;
                herm_pair(u[bitrev(q, LC)], v[bitrev((C - q) & (C - 1), LC)], q ? cmul(wns, W[16 * q]) : wns, xk, xmk); // w_n^(M q) = w_{2C}^q
                st_stream(out_ptr(row, k + UM * (unsigned) q), xk);
                if (!(last_single && q == nq - 1)) st_stream(out_ptr(row, (UM - k) + UM * (unsigned) (C - 1 - q)), xmk);
            }
        }
    };
    auto rev4 = [](int s) { return (int) (__brev((unsigned) s) >> 28); };

    if (blockIdx.x < nxl) stage(V + (unsigned long long) blockIdx.x * pitch, 0, keep);

    for (unsigned row = blockIdx.x; row < nxl; row += gridDim.x) {
#pragma unroll 1
        for (int c = 0; c < C; ++c) {
            cp_async_wait_all();
            __syncthreads(); // (1) the sub-sequence has landed and is visible to every warp
            rd2::passes_ab(sm, twA, warp, lane);
            __syncthreads(); // (2) all 16 sub-spectra are complete
            cd A[16], B[16];
            rd2::load_columns(sm, lt, A, B);
            __syncthreads(); // (3) the pencil is dead: refill it with the next class / the first class of the next row
            if (c < C - 2)
                stage(V + (unsigned long long) row * pitch, c + 1, keep);
            else if (c == C - 2)
                stage(V + (unsigned long long) row * pitch, C - 1, drop);
            else if (row + gridDim.x < nxl)
                stage(V + (unsigned long long) (row + gridDim.x) * pitch, 0, keep);
            rd2::finish_columns(tw2, lt, A, B);
            // park in butterfly order: slot i <- A[i], slot 16 + i <- B[i]
            cd *xc = xe + (size_t) c * 32 * ROW_THREADS;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                st_cg(xc + i * ROW_THREADS, A[i]);
                st_cg(xc + (16 + i) * ROW_THREADS, B[i]);
            }
        }

        // ---- combine (the gather of the next row is in flight) ----
        cd u0[C], v0[C];
        if (lt != 0) {
            // k = lt + s PP (register bitrev(s) of column jA) pairs with M - k = jB + (15 - s) PP (register bitrev((16 - s) & 15) of jB)
            cd T[C];
            T[0] = make_double2(1.0, 0.0);
#pragma unroll
            for (int c = 1; c < C; ++c) T[c] = ldtw(tw, 2u * (unsigned) c * (unsigned) lt);
            const cd wnj = tw3n[lt];
            cd u1[C], v1[C];
            load_pair(u0, v0, rev4(0), 16 + rev4(0));
#pragma unroll 1
            for (int s = 0; s < 16; s += 2) {
                load_pair(u1, v1, rev4(s + 1), 16 + rev4(15 - s));
                emit(row, u0, v0, (unsigned) lt, s, T, wnj, C, false);
                if (s + 2 < 16) load_pair(u0, v0, rev4(s + 2), 16 + rev4(14 - s));
                emit(row, u1, v1, (unsigned) lt, s + 1, T, wnj, C, false);
            }
        } else {
            // columns 0 and PP/2 are their own partners; natural output s sits in register bitrev(s) of both
            cd T[C];
#pragma unroll
            for (int c = 0; c < C; ++c) T[c] = make_double2(1.0, 0.0);
            const cd one = make_double2(1.0, 0.0);
            load_pair(u0, v0, 0, 0); // k = 0: bins M q; q = 0 gives X[0] and X[m], q = C/2 is its own partner
            emit(row, u0, v0, 0u, 0, T, one, C / 2 + 1, true);
#pragma unroll 1
            for (int s = 1; s < 8; ++s) {
                load_pair(u0, v0, rev4(s), rev4(16 - s));
                emit(row, u0, v0, 0u, s, T, one, C, false);
            }
            load_pair(u0, v0, rev4(8), rev4(8)); // k = M/2 is its own partner
            emit(row, u0, v0, 0u, 8, T, one, C / 2, false);
#pragma unroll
            for (int c = 1; c < C; ++c) T[c] = ldtw(tw, (unsigned) c * (unsigned) PP); // w_m^(c PP/2)
            const cd wnh = tw3n[PP / 2];
#pragma unroll 1
            for (int s = 0; s < 8; ++s) {
                load_pair(u0, v0, 16 + rev4(s), 16 + rev4(15 - s));
                emit(row, u0, v0, (unsigned) (PP / 2), s, T, wnh, C, false);
            }
        }
    }
}

}  // namespace hpxfft_b200

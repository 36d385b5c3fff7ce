// launch_rows.cu -- dispatch of the row (r2c) kernels.
#include "kernels_rows_long.cuh"
#include "kernels_rows_long2.cuh"
#include "kernels_rows_dit2.cuh"
#include "kernels_rows_ditc.cuh"
#include "kernels_rows_v2.cuh"

#include <cstdlib>
#include "launch_util.h"

namespace hpxfft_b200 {

namespace {

template <int M, bool FAST> int launch_rows_big_t(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch)
{
    constexpr size_t smem = row_smem_total<M>();
    if (int rc = ensure_smem(rows_r2c_kernel<M, FAST>, smem, p->device)) return rc;
    const unsigned ngroups = (nrows + row_group<M>() - 1) / row_group<M>();
    // one resident CTA per SM (shared memory bound): persistent CTAs amortise the twiddle-table build
    const int sms = p->sm_count - p->sm_reserve;
    const unsigned cap = sms > 0 ? (unsigned) sms : 1u;
    const unsigned grid = ngroups < cap ? ngroups : cap;
    rows_r2c_kernel<M, FAST><<<grid, ROW_THREADS, smem, p->stream>>>(V, pitch, nrows, dst, p->tw_row);
    CU(cudaGetLastError());
    return 0;
}

template <int M> int launch_rows_big(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch)
{
    // fast output addressing: one destination rank and tile-aligned per-s stride (the 1-GPU hot config)
    if constexpr (M == 8192) {
        if (dst.P == 1) return launch_rows_big_t<M, true>(p, dst, nrows, V, pitch);
    }
    return launch_rows_big_t<M, false>(p, dst, nrows, V, pitch);
}

// rows longer than one pencil: one persistent CTA per row, C sequential sub-FFTs, L2-resident scratch (kernels_rows_long.cuh)
template <int C> int launch_rows_long(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch)
{
    constexpr size_t smem = rows_long_smem_bytes<C>();
    if (int rc = ensure_smem(rows_long_kernel<C>, smem, p->device)) return rc;
    if (!p->zraw) return fail(HPXFFT_B200_ESTATE, "long-row scratch missing");
    const unsigned cap = (unsigned) (p->sm_count - p->sm_reserve > 0 ? p->sm_count - p->sm_reserve : 1);
    const unsigned grid = nrows < cap ? nrows : cap;
    rows_long_kernel<C><<<grid, ROW_THREADS, smem, p->stream>>>(V, pitch, nrows, dst, p->tw_row, p->zraw);
    CU(cudaGetLastError());
    return 0;
}

// ny = 16384: warp-local in-place sub-FFTs (kernels_rows_v2.cuh); HPXFFT_B200_ROWS_V1=1 selects the Stockham kernel (A/B runs)
template <bool FAST, bool PF, bool ILV = false> int launch_rows_v2_t(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch)
{
    if (int rc = ensure_smem(rows_r2c_v2_kernel<FAST, PF, ILV>, rv2::SMEM, p->device)) return rc;
    const unsigned cap = (unsigned) (p->sm_count - p->sm_reserve > 0 ? p->sm_count - p->sm_reserve : 1);
    const unsigned grid = nrows < cap ? nrows : cap;
    rows_r2c_v2_kernel<FAST, PF, ILV><<<grid, ROW_THREADS, rv2::SMEM, p->stream>>>(V, pitch, nrows, dst, p->tw_row);
    CU(cudaGetLastError());
    return 0;
}

// ny = 32768: both DIF halves in one CTA, even bins parked in L2, 256-bit paired stores (kernels_rows_long2.cuh)
template <bool FAST> int launch_rows_long2_t(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch)
{
    if (int rc = ensure_smem(rows_long2_kernel<FAST>, rl2::SMEM, p->device)) return rc;
    if (!p->zraw) return fail(HPXFFT_B200_ESTATE, "long-row scratch missing");
    const unsigned cap = (unsigned) (p->sm_count - p->sm_reserve > 0 ? p->sm_count - p->sm_reserve : 1);
    const unsigned grid = nrows < cap ? nrows : cap;
    rows_long2_kernel<FAST><<<grid, ROW_THREADS, rl2::SMEM, p->stream>>>(V, pitch, nrows, dst, p->tw_row, p->zraw);
    CU(cudaGetLastError());
    return 0;
}

// ny = 32768, decimation in time: two v2-style halves by sample parity, Ze parked per thread in L2 (kernels_rows_dit2.cuh)
template <bool FAST, bool PF> int launch_rows_dit2_t(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch)
{
    if (int rc = ensure_smem(rows_dit2_kernel<FAST, PF>, rd2::SMEM, p->device)) return rc;
    if (!p->zraw) return fail(HPXFFT_B200_ESTATE, "long-row scratch missing");
    const unsigned cap = (unsigned) (p->sm_count - p->sm_reserve > 0 ? p->sm_count - p->sm_reserve : 1);
    const unsigned grid = nrows < cap ? nrows : cap;
    rows_dit2_kernel<FAST, PF><<<grid, ROW_THREADS, rd2::SMEM, p->stream>>>(V, pitch, nrows, dst, p->tw_row, p->zraw);
    CU(cudaGetLastError());
    return 0;
}

// ny = 2 C * 8192, decimation in time over C sample classes, all classes parked per thread (kernels_rows_ditc.cuh)
template <int C, bool FAST> int launch_rows_ditc_t(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch)
{
    constexpr size_t smem = rdc::smem_bytes<C>();
    if (int rc = ensure_smem(rows_ditc_kernel<C, FAST>, smem, p->device)) return rc;
    if (!p->zraw) return fail(HPXFFT_B200_ESTATE, "long-row scratch missing");
    const unsigned cap = (unsigned) (p->sm_count - p->sm_reserve > 0 ? p->sm_count - p->sm_reserve : 1);
    const unsigned grid = nrows < cap ? nrows : cap;
    rows_ditc_kernel<C, FAST><<<grid, ROW_THREADS, smem, p->stream>>>(V, pitch, nrows, dst, p->tw_row, p->zraw);
    CU(cudaGetLastError());
    return 0;
}
template <int C> int launch_rows_ditc(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch, bool general)
{
    return dst.P == 1 && !general ? launch_rows_ditc_t<C, true>(p, dst, nrows, V, pitch) : launch_rows_ditc_t<C, false>(p, dst, nrows, V, pitch);
}

// Environment knobs, read per launch (cheap) so that the parity tests can select the variants in one process.
// HPXFFT_B200_ROWS_LONG selects the kernel for rows longer than one pencil (ny >= 32768):
//   unset / 0: default -- rows_dit2_kernel (ny = 32768), rows_ditc_kernel<4 | 8> (ny = 65536 | 131072)
//   1: rows_long_kernel<C> (round-1 design: C passes over the row, assembly loop)      2: rows_long2_kernel (ny = 32768 only)
//   3: rows_dit2_kernel (ny = 32768 only)                                                5: rows_ditc_kernel<C>
int rows_long_variant()
{
    const char *e = getenv("HPXFFT_B200_ROWS_LONG");
    return e ? atoi(e) : 0;
}
// HPXFFT_B200_ROWS_GENERAL=1: the general output addressing (the one the distributed slabs use) on one GPU -- for the tests
bool rows_general()
{
    const char *e = getenv("HPXFFT_B200_ROWS_GENERAL");
    return e && e[0] == '1';
}
// HPXFFT_B200_ROWS_PF=0|1: bulk L2 prefetch of the next row.  Default: on for ny = 32768 (6.76 vs 7.10 ms at 32768^2), off for
// ny = 16384 (1.170 vs 1.153 ms at 16384^2).
bool rows_prefetch(bool dflt)
{
    const char *e = getenv("HPXFFT_B200_ROWS_PF");
    return e ? e[0] == '1' : dflt;
}

// HPXFFT_B200_ROWS_ILV=0|1: ny = 16384, refill of the pencil issued in groups between the steps of the tail instead of one burst.
// Default: on for the one-GPU addressing (rows 1.129 vs 1.137 ms at 16384^2, A/B/A/B in profiles/r2_x_summary.txt), off otherwise
// (not measured with several destination ranks).
bool rows_interleaved(bool dflt)
{
    const char *e = getenv("HPXFFT_B200_ROWS_ILV");
    return e ? e[0] == '1' : dflt;
}

bool rows_v1_path()
{
    const char *e = getenv("HPXFFT_B200_ROWS_V1");
    return e && e[0] == '1';
}

template <int M> int launch_rows_tiny(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch)
{
    const unsigned block = 128, grid = (nrows + block - 1) / block;
    rows_r2c_tiny_kernel<M><<<grid, block, 0, p->stream>>>(V, pitch, nrows, dst, p->tw_row);
    CU(cudaGetLastError());
    return 0;
}

}  // namespace

int rows_launch_count(size_t) { return 1; }

int launch_rows(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch, size_t m)
{
    if (p->rows_mixed) return launch_rows_mixed(p, dst, nrows, V, pitch, m);
    if (p->rows_blue) return launch_rows_blue(p, dst, nrows, V, pitch);
    if (p->rows_generic) return launch_rows_generic(p, dst, nrows, V, pitch, m);
    // Several destination ranks (P > 1).
    // ny = 32768: the general-addressing instantiation rows_dit2_kernel<false> runs at half the rate of <true> -- 12.7 against
    // 6.8 ms for 32768 rows on one GPU (profiles/r2_w_bench_32768_general_v0.json), 6.97 / 7.09 ms per 16384-row slab on 2 GPUs
    // over the fused / copy-engine transports -- while rows_long2_kernel<false> loses 5 % (7.9 ms, 3.6 ms per slab).  Slabs with
    // several destination ranks therefore keep rows_long2_kernel.
    // ny = 65536 / 131072: rows_ditc_kernel<C, false> loses 3 % (3.82 against 3.70 ms, 2048 x 131072) and stays the default
    // unless its stores cross NVLink (fused transport): its mirrored bin families start one bin off a 512-byte boundary, every
    // warp store leaves a 16-byte straggler that L2 merges for local stores and a peer window does not, and that case has not
    // been measured -- it keeps rows_long_kernel<C>, whose mirrored warp stores are aligned.
    // HPXFFT_B200_ROWS_LONG=3 / 5 forces the decimation-in-time kernels anyway (the distributed parity tests do).
    const bool multi = dst.P > 1, remote = multi && p->transport == TR_FUSED;
    switch (m) {
    case 1: return launch_rows_tiny<1>(p, dst, nrows, V, pitch);
    case 2: return launch_rows_tiny<2>(p, dst, nrows, V, pitch);
    case 4: return launch_rows_tiny<4>(p, dst, nrows, V, pitch);
    case 8: return launch_rows_tiny<8>(p, dst, nrows, V, pitch);
    case 16: return launch_rows_tiny<16>(p, dst, nrows, V, pitch);
    case 32: return launch_rows_big<32>(p, dst, nrows, V, pitch);
    case 64: return launch_rows_big<64>(p, dst, nrows, V, pitch);
    case 128: return launch_rows_big<128>(p, dst, nrows, V, pitch);
    case 256: return launch_rows_big<256>(p, dst, nrows, V, pitch);
    case 512: return launch_rows_big<512>(p, dst, nrows, V, pitch);
    case 1024: return launch_rows_big<1024>(p, dst, nrows, V, pitch);
    case 2048: return launch_rows_big<2048>(p, dst, nrows, V, pitch);
    case 4096: return launch_rows_big<4096>(p, dst, nrows, V, pitch);
    case 8192: {
        if (rows_v1_path()) return launch_rows_big<8192>(p, dst, nrows, V, pitch);
        const bool fast = dst.P == 1 && !rows_general();
        if (rows_interleaved(fast) && !rows_prefetch(false)) return fast ? launch_rows_v2_t<true, false, true>(p, dst, nrows, V, pitch) : launch_rows_v2_t<false, false, true>(p, dst, nrows, V, pitch);
        if (rows_prefetch(false)) return fast ? launch_rows_v2_t<true, true>(p, dst, nrows, V, pitch) : launch_rows_v2_t<false, true>(p, dst, nrows, V, pitch);
        return fast ? launch_rows_v2_t<true, false>(p, dst, nrows, V, pitch) : launch_rows_v2_t<false, false>(p, dst, nrows, V, pitch);
    }
    case 16384: {
        const int v = rows_long_variant();
        const bool fast = dst.P == 1 && !rows_general();
        if (v == 1) return launch_rows_long<2>(p, dst, nrows, V, pitch);
        if (v == 2 || (v == 0 && multi)) return fast ? launch_rows_long2_t<true>(p, dst, nrows, V, pitch) : launch_rows_long2_t<false>(p, dst, nrows, V, pitch);
        if (v == 5) return launch_rows_ditc<2>(p, dst, nrows, V, pitch, !fast);
        if (rows_prefetch(true)) return fast ? launch_rows_dit2_t<true, true>(p, dst, nrows, V, pitch) : launch_rows_dit2_t<false, true>(p, dst, nrows, V, pitch);
        return fast ? launch_rows_dit2_t<true, false>(p, dst, nrows, V, pitch) : launch_rows_dit2_t<false, false>(p, dst, nrows, V, pitch);
    }
    case 32768:
        if (rows_long_variant() == 1 || (rows_long_variant() == 0 && remote)) return launch_rows_long<4>(p, dst, nrows, V, pitch);
        return launch_rows_ditc<4>(p, dst, nrows, V, pitch, rows_general());
    case 65536:
        if (rows_long_variant() == 1 || (rows_long_variant() == 0 && remote)) return launch_rows_long<8>(p, dst, nrows, V, pitch);
        return launch_rows_ditc<8>(p, dst, nrows, V, pitch, rows_general());
    default: return fail(HPXFFT_B200_EINVAL, "unsupported row length ny=%zu (ny/2 must be a power of two <= 65536)", 2 * m);
    }
}

}  // namespace hpxfft_b200

// launch_rows.cu -- dispatch of the row (r2c) kernels; rows of up to one shared-memory pencil (ny <= 16384).
// Longer rows: launch_rows_long.cu (dispatch, DIF kernels, rows_dit2_kernel) and launch_rows_ditc.cu.
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

}  // namespace

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

namespace {

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
    case 16384:
    case 32768:
    case 65536: return launch_rows_longer(p, dst, nrows, V, pitch, m);
    default: return fail(HPXFFT_B200_EINVAL, "unsupported row length ny=%zu (ny/2 must be a power of two <= 65536)", 2 * m);
    }
}

}  // namespace hpxfft_b200

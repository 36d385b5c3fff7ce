// launch_rows_long.cu -- rows longer than one shared-memory pencil (ny = 32768, 65536, 131072): dispatch, the
// decimation-in-frequency kernels (rows_long_kernel<C>, rows_long2_kernel) and rows_dit2_kernel.  rows_ditc_kernel<C> is compiled
// in launch_rows_ditc.cu (separate translation units keep the build parallel).
#include "kernels_rows_long.cuh"
#include "kernels_rows_long2.cuh"
#include "kernels_rows_dit2.cuh"

#include <cstdlib>
#include "launch_util.h"

namespace hpxfft_b200 {

namespace {

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
}  // namespace

int launch_rows_longer(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch, size_t m)
{
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
    case 16384: {
        const int v = rows_long_variant();
        const bool fast = dst.P == 1 && !rows_general();
        if (v == 1) return launch_rows_long<2>(p, dst, nrows, V, pitch);
        if (v == 2 || (v == 0 && multi)) return fast ? launch_rows_long2_t<true>(p, dst, nrows, V, pitch) : launch_rows_long2_t<false>(p, dst, nrows, V, pitch);
        if (v == 5) return launch_rows_ditc(p, dst, nrows, V, pitch, 2, !fast);
        if (rows_prefetch(true)) return fast ? launch_rows_dit2_t<true, true>(p, dst, nrows, V, pitch) : launch_rows_dit2_t<false, true>(p, dst, nrows, V, pitch);
        return fast ? launch_rows_dit2_t<true, false>(p, dst, nrows, V, pitch) : launch_rows_dit2_t<false, false>(p, dst, nrows, V, pitch);
    }
    case 32768:
        if (rows_long_variant() == 1 || (rows_long_variant() == 0 && remote)) return launch_rows_long<4>(p, dst, nrows, V, pitch);
        return launch_rows_ditc(p, dst, nrows, V, pitch, 4, rows_general());
    case 65536:
        if (rows_long_variant() == 1 || (rows_long_variant() == 0 && remote)) return launch_rows_long<8>(p, dst, nrows, V, pitch);
        return launch_rows_ditc(p, dst, nrows, V, pitch, 8, rows_general());
    default: return fail(HPXFFT_B200_EINVAL, "unsupported row length ny=%zu", 2 * m);
    }
}

}  // namespace hpxfft_b200

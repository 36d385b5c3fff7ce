// launch_rows_ditc.cu -- rows_ditc_kernel<C> (kernels_rows_ditc.cuh): decimation in time over C = 2, 4, 8 sample classes.
#include "kernels_rows_ditc.cuh"

#include <cstdlib>
#include "launch_util.h"

namespace hpxfft_b200 {

namespace {

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
template <int C> int launch_rows_ditc_c(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch, bool general)
{
    return dst.P == 1 && !general ? launch_rows_ditc_t<C, true>(p, dst, nrows, V, pitch) : launch_rows_ditc_t<C, false>(p, dst, nrows, V, pitch);
}

}  // namespace

int launch_rows_ditc(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch, int C, bool general)
{
    switch (C) {
    case 2: return launch_rows_ditc_c<2>(p, dst, nrows, V, pitch, general);
    case 4: return launch_rows_ditc_c<4>(p, dst, nrows, V, pitch, general);
    case 8: return launch_rows_ditc_c<8>(p, dst, nrows, V, pitch, general);
    default: return fail(HPXFFT_B200_EINVAL, "rows_ditc_kernel: C = %d", C);
    }
}

}  // namespace hpxfft_b200

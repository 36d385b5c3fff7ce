// launch_misc.cu -- synthetic fill, exchange #2 unpack, layout helpers of the adapter-level entry points,
// fused-pair group dispatch.
#include "kernels_misc.cuh"
#include "launch_util.h"

namespace hpxfft_b200 {

#define DECL_GROUP(g)                                                                                                          \
    int launch_cols_fused_g##g(const hpxfft_b200_plan *, const InterView &, const ColDst &, unsigned, unsigned, int *);        \
    int fused_blocks_per_sm_g##g(unsigned, unsigned, unsigned, int *, int *);
DECL_GROUP(0) DECL_GROUP(1) DECL_GROUP(2) DECL_GROUP(3)
#undef DECL_GROUP

int launch_cols_fused(const hpxfft_b200_plan *p, const InterView &in, const ColDst &out, unsigned ct0, unsigned ntiles)
{
    if (ntiles == 0) return 0;
    int rc = 0;
    if (!launch_cols_fused_g0(p, in, out, ct0, ntiles, &rc)) return rc;
    if (!launch_cols_fused_g1(p, in, out, ct0, ntiles, &rc)) return rc;
    if (!launch_cols_fused_g2(p, in, out, ct0, ntiles, &rc)) return rc;
    if (!launch_cols_fused_g3(p, in, out, ct0, ntiles, &rc)) return rc;
    return fail(HPXFFT_B200_EINVAL, "no fused column kernel for %u x %u", p->n1, p->n2);
}

int fused_blocks_per_sm(unsigned n1, unsigned n2, unsigned split, int *bps)
{
    int rc = 0;
    if (!fused_blocks_per_sm_g0(n1, n2, split, bps, &rc)) return rc;
    if (!fused_blocks_per_sm_g1(n1, n2, split, bps, &rc)) return rc;
    if (!fused_blocks_per_sm_g2(n1, n2, split, bps, &rc)) return rc;
    if (!fused_blocks_per_sm_g3(n1, n2, split, bps, &rc)) return rc;
    return fail(HPXFFT_B200_EINVAL, "no fused column kernel for %u x %u", n1, n2);
}

bool fused_pair_exists(unsigned n1, unsigned n2)
{
    return n1 >= 32 && n1 <= 512 && (n2 == n1 || n2 * 2 == n1) && n2 >= 16;
}

int launch_fill(const hpxfft_b200_plan *p, int pattern, unsigned long long seed)
{
    fill_kernel<<<(unsigned) p->sm_count * 8, 256, 0, p->stream>>>(p->V, (unsigned) p->nxl, (unsigned) p->ny, (unsigned) p->n_col,
                                                                    (unsigned long long) p->rank * p->nxl, pattern, seed);
    CU(cudaGetLastError());
    return 0;
}

int launch_unpack(const hpxfft_b200_plan *p, cudaStream_t s)
{
    unpack_kernel<<<dim3((unsigned) p->nxl, (unsigned) p->P), 256, 0, s>>>(p->bufB, (cd *) p->V, (unsigned) p->nxl, (unsigned) p->cy, p->wq0,
                                                                           (unsigned) p->P, (unsigned) p->rank);
    CU(cudaGetLastError());
    return 0;
}

int launch_unpack_chunk(const hpxfft_b200_plan *p, const UnpackChunk &u, cudaStream_t s)
{
    unpack_chunk_kernel<<<dim3((unsigned) p->nxl, (unsigned) p->P), 256, 0, s>>>(p->bufC, (cd *) p->V, (unsigned) p->cy, u);
    CU(cudaGetLastError());
    return 0;
}

int launch_tile(const cd *A, cd *I, unsigned n, unsigned width, cudaStream_t s)
{
    tile_kernel<<<148 * 4, 256, 0, s>>>(A, I, n, width);
    CU(cudaGetLastError());
    return 0;
}

int launch_untile(const cd *I, cd *A, unsigned n, unsigned width, cudaStream_t s)
{
    untile_kernel<<<148 * 4, 256, 0, s>>>(I, A, n, width);
    CU(cudaGetLastError());
    return 0;
}

int launch_rows_generic(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch, size_t m)
{
    const unsigned ny = (unsigned) (2 * m), cy = ny / 2 + 1;
    rows_generic_kernel<<<dim3(nrows, (cy + 127) / 128), 128, 0, p->stream>>>((const double *) V, 2 * pitch, nrows, ny, dst, p->tw_row);
    CU(cudaGetLastError());
    return 0;
}

int launch_cols_generic(const hpxfft_b200_plan *p, const InterView &in, const ColDst &out, unsigned ntiles, unsigned nx)
{
    cols_generic_kernel<<<dim3(ntiles * CW, (nx + 127) / 128), 128, 0, p->stream>>>(in, out, nx, p->tw_col);
    CU(cudaGetLastError());
    return 0;
}

}  // namespace hpxfft_b200

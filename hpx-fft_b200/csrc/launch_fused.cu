// launch_fused.cu -- dispatch of the persistent fused four-step column kernel.  Compiled once per group of
// (N1, N2) pairs (-DHPXFFT_B200_FUSED_GROUP=g) so that the instantiations build in parallel.
#include "kernels_cols.cuh"
#include "launch_util.h"

#include <cstdlib>

#ifndef HPXFFT_B200_FUSED_GROUP
#error "compile with -DHPXFFT_B200_FUSED_GROUP=0..3"
#endif

#if HPXFFT_B200_FUSED_GROUP == 0
#define FUSED_PAIRS(X) X(32, 16) X(32, 32) X(64, 32) X(64, 64)
#define GROUP_FN(name) name##_g0
#elif HPXFFT_B200_FUSED_GROUP == 1
#define FUSED_PAIRS(X) X(128, 64) X(128, 128)
#define GROUP_FN(name) name##_g1
#elif HPXFFT_B200_FUSED_GROUP == 2
#define FUSED_PAIRS(X) X(256, 128) X(256, 256)
#define GROUP_FN(name) name##_g2
#else
#define FUSED_PAIRS(X) X(512, 256) X(512, 512)
#define GROUP_FN(name) name##_g3
#endif

namespace hpxfft_b200 {

namespace {

template <int N1, int N2, int SPLIT> int fused_occupancy(int *blocks_per_sm)
{
    constexpr size_t smem = fused_smem_bytes<N1, N2, SPLIT>();
    if (int rc = set_smem(cols_fused_kernel<N1, N2, SPLIT>, smem)) return rc;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, cols_fused_kernel<N1, N2, SPLIT>, fused_threads<N1, N2>(), smem));
    return 0;
}

// ct0 / nstrips in real strips; the kernel sees SPLIT virtual strips per real one
template <int N1, int N2, int SPLIT>
int launch_cols_fused_t(const hpxfft_b200_plan *p, const InterView &in, const ColDst &out, unsigned ct0, unsigned nstrips)
{
    constexpr size_t smem = fused_smem_bytes<N1, N2, SPLIT>();
    if (int rc = ensure_smem(cols_fused_kernel<N1, N2, SPLIT>, smem, p->device)) return rc;
    const unsigned ntiles = nstrips * SPLIT;
    CU(cudaMemsetAsync(p->ctl, 0, (1 + 2 * (size_t) ntiles) * sizeof(unsigned), p->stream));
    FusedCtl ctl;
    ctl.counter = p->ctl;
    ctl.doneA = p->ctl + 1;
    ctl.doneB = p->ctl + 1 + ntiles;
    ctl.lag = p->lag;
    ctl.nslot = p->nslot;
    ctl.ct0 = ct0;
    if (ctl.nslot > ntiles) ctl.nslot = ntiles; // a short chunk needs (and may use) no more slots than strips
    {
        static const int discard = [] {
            const char *e = getenv("HPXFFT_B200_DISCARD");
            return (e && e[0] == '0') ? 0 : 1;
        }();
        // peer destinations: keep the level-B signal fence-free (kernels_cols.cuh) at the price of writing the dead scratch lines back
        ctl.discard = (p->transport == TR_FUSED && p->P > 1) ? 0u : (unsigned) discard;
    }
    cols_fused_kernel<N1, N2, SPLIT><<<p->fused_grid, fused_threads<N1, N2>(), smem, p->stream>>>(in, p->S, out, p->tw_col, p->tw_il, ntiles, ctl);
    CU(cudaGetLastError());
    return 0;
}

}  // namespace

// returns 1 when the pair does not belong to this group
int GROUP_FN(launch_cols_fused)(const hpxfft_b200_plan *p, const InterView &in, const ColDst &out, unsigned ct0, unsigned ntiles, int *rc)
{
#define X(A, B) if (p->n1 == A && p->n2 == B && p->col_split == 1) { *rc = launch_cols_fused_t<A, B, 1>(p, in, out, ct0, ntiles); return 0; }
    FUSED_PAIRS(X)
#undef X
#if HPXFFT_B200_FUSED_GROUP == 1
    if (p->n1 == 128 && p->n2 == 128 && p->col_split == 2) { *rc = launch_cols_fused_t<128, 128, 2>(p, in, out, ct0, ntiles); return 0; }
#endif
    return 1;
}

int GROUP_FN(fused_blocks_per_sm)(unsigned n1, unsigned n2, unsigned split, int *bps, int *rc)
{
#define X(A, B) if (n1 == A && n2 == B && split == 1) { *rc = fused_occupancy<A, B, 1>(bps); return 0; }
    FUSED_PAIRS(X)
#undef X
#if HPXFFT_B200_FUSED_GROUP == 1
    if (n1 == 128 && n2 == 128 && split == 2) { *rc = fused_occupancy<128, 128, 2>(bps); return 0; }
#endif
    return 1;
}

}  // namespace hpxfft_b200

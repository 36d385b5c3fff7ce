// launch_cols.cu -- dispatch of the unfused column (c2c) kernels: single tile FFT, four-step level A / level B.
#include "kernels_cols.cuh"
#include "launch_util.h"

namespace hpxfft_b200 {

namespace {

template <int N> int launch_cols_single(const hpxfft_b200_plan *p, const InterView &in, const ColDst &out, unsigned ntiles)
{
    constexpr size_t smem = single_smem_bytes<N>();
    if (int rc = ensure_smem(cols_single_kernel<N>, smem, p->device)) return rc;
    cols_single_kernel<N><<<ntiles, col_threads(N), smem, p->stream>>>(in, out, p->tw_col);
    CU(cudaGetLastError());
    return 0;
}

template <int N1> int launch_cols_A(const hpxfft_b200_plan *p, const InterView &in, cd *S, unsigned n2, unsigned ntiles)
{
    constexpr size_t smem = levelA_smem_bytes<N1>();
    if (int rc = ensure_smem(cols_levelA_kernel<N1>, smem, p->device)) return rc;
    cols_levelA_kernel<N1><<<dim3(n2, ntiles), col_threads(N1), smem, p->stream>>>(in, S, n2, p->tw_col, p->tw_il);
    CU(cudaGetLastError());
    return 0;
}

template <int N2> int launch_cols_B(const hpxfft_b200_plan *p, const cd *S, const ColDst &out, unsigned n1, unsigned ntiles)
{
    constexpr size_t smem = single_smem_bytes<N2>();
    if (int rc = ensure_smem(cols_levelB_kernel<N2>, smem, p->device)) return rc;
    cols_levelB_kernel<N2><<<dim3(n1, ntiles), col_threads(N2), smem, p->stream>>>(S, out, n1, p->tw_col);
    CU(cudaGetLastError());
    return 0;
}

#define DISPATCH_POW2(FN, N, LO, ...)                                                         \
    switch (N) {                                                                              \
    case 1: if (LO <= 1) return FN<1>(__VA_ARGS__); break;                                    \
    case 2: if (LO <= 2) return FN<2>(__VA_ARGS__); break;                                    \
    case 4: if (LO <= 4) return FN<4>(__VA_ARGS__); break;                                    \
    case 8: if (LO <= 8) return FN<8>(__VA_ARGS__); break;                                    \
    case 16: return FN<16>(__VA_ARGS__);                                                      \
    case 32: return FN<32>(__VA_ARGS__);                                                      \
    case 64: return FN<64>(__VA_ARGS__);                                                      \
    case 128: return FN<128>(__VA_ARGS__);                                                    \
    case 256: return FN<256>(__VA_ARGS__);                                                    \
    case 512: return FN<512>(__VA_ARGS__);                                                    \
    default: break;                                                                           \
    }

}  // namespace

int launch_cols(const hpxfft_b200_plan *p, const InterView &in, const ColDst &out, unsigned ntiles, cd *S, unsigned nx, unsigned n1,
                unsigned n2, bool two_level, int *launches, cudaEvent_t mid)
{
    if (p->cols_mixed) {
        if (mid) CU(cudaEventRecord(mid, p->stream));
        return launch_cols_mixed(p, in, out, launches);
    }
    if (p->cols_blue) {
        if (mid) CU(cudaEventRecord(mid, p->stream));
        return launch_cols_blue(p, in, out, launches);
    }
    if (p->cols_generic) {
        if (launches) *launches += 1;
        if (mid) CU(cudaEventRecord(mid, p->stream));
        return launch_cols_generic(p, in, out, ntiles, nx);
    }
    if (!two_level) {
        if (launches) *launches += 1;
        if (mid) CU(cudaEventRecord(mid, p->stream));
        if (nx <= 256) { DISPATCH_POW2(launch_cols_single, nx, 1, p, in, out, ntiles) }
        return fail(HPXFFT_B200_EINVAL, "unsupported single-level column length %u", nx);
    }
    if (launches) *launches += 2;
    {
        auto a = [&]() -> int {
            DISPATCH_POW2(launch_cols_A, n1, 16, p, in, S, n2, ntiles)
            return fail(HPXFFT_B200_EINVAL, "unsupported level-A length %u", n1);
        };
        if (int rc = a()) return rc;
        if (mid) CU(cudaEventRecord(mid, p->stream));
    }
    DISPATCH_POW2(launch_cols_B, n2, 16, p, S, out, n1, ntiles)
    return fail(HPXFFT_B200_EINVAL, "unsupported level-B length %u", n2);
}

}  // namespace hpxfft_b200

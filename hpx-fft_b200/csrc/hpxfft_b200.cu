// hpxfft_b200.cu -- plan object and C ABI (include/hpxfft_b200.h) of libhpxfft_b200.so.
//
// Host-side counterpart of hpxfft::shared::loop::initialize / fft_2d_r2c_par
// (core/src/shared/loop.cpp:56-113,158-189) and hpxfft::distributed::loop::initialize / fft_2d_r2c
// (core/src/distributed/loop.cpp:130-347) of the reference: dimension inference, buffers, "plans"
// (twiddle tables + kernel selection), communicator, the phase sequence and its timers.
#include "../../include/hpxfft_b200.h"

#include "kernels_cols.cuh"
#include "kernels_misc.cuh"
#include "kernels_rows.cuh"
#include "kernels_rows16.cuh"

#include <nccl.h>  // types only: the library is dlopen'ed lazily (see NcclApi) so that a host process
                   // that already carries its own libnccl.so.2 (e.g. PyTorch's bundled one) is reused
#include <dlfcn.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace hpxfft_b200;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                            \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess)                                                                              \
            return fail(HPXFFT_B200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                                          \
    } while (0)

#define NC(call)                                                                                            \
    do {                                                                                                    \
        ncclResult_t r_ = (call);                                                                           \
        if (r_ != ncclSuccess)                                                                              \
            return fail(HPXFFT_B200_ENCCL, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, \
                        __LINE__);                                                                          \
    } while (0)

// NCCL entry points, resolved at first use.
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitRankConfig)(ncclComm_t *, int, ncclUniqueId, int, ncclConfig_t *) = nullptr; // optional
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
};
NcclApi g_nccl;

int nccl_load()
{
    if (g_nccl.handle) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return fail(HPXFFT_B200_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
#define NSYM(field, name)                                                                      \
    *(void **) (&g_nccl.field) = dlsym(h, name);                                               \
    if (!g_nccl.field) return fail(HPXFFT_B200_ENCCL, "libnccl: missing symbol %s", name);
    NSYM(GetUniqueId, "ncclGetUniqueId")
    NSYM(CommInitRank, "ncclCommInitRank")
    NSYM(CommDestroy, "ncclCommDestroy")
    NSYM(GetErrorString, "ncclGetErrorString")
    NSYM(GroupStart, "ncclGroupStart")
    NSYM(GroupEnd, "ncclGroupEnd")
    NSYM(Send, "ncclSend")
    NSYM(Recv, "ncclRecv")
    NSYM(AllReduce, "ncclAllReduce")
#undef NSYM
    *(void **) (&g_nccl.CommInitRankConfig) = dlsym(h, "ncclCommInitRankConfig");
    g_nccl.handle = h;
    return 0;
}

enum Mode { MODE_SHARED = 0, MODE_SCATTER = 1, MODE_ALL_TO_ALL = 2, MODE_P2P = 3 };

bool is_pow2(size_t v) { return v && !(v & (v - 1)); }

// exp(-2 pi i k / n) rounded from long double; exact on the axes and diagonals so that small
// integer inputs (the reference's 4x4 known-answer test) transform exactly.
void make_twiddles(std::vector<double2> &t, size_t n)
{
    t.resize(n);
    const long double PI_L = 3.14159265358979323846264338327950288L;
    for (size_t k = 0; k < n; ++k) {
        // reduce to the first octant, evaluate there, map back by symmetry
        const size_t k8 = (8 * k) / n;           // octant 0..7
        const bool on_oct = (8 * k) % n == 0;
        long double c, s; // cos, sin of 2 pi k / n
        if (on_oct) {
            static const long double r2 = 0.70710678118654752440084436210484903928L;
            const long double C[8] = {1, r2, 0, -r2, -1, -r2, 0, r2};
            const long double S[8] = {0, r2, 1, r2, 0, -r2, -1, -r2};
            c = C[k8];
            s = S[k8];
        } else {
            const long double a = 2.0L * PI_L * (long double) k / (long double) n;
            c = cosl(a);
            s = sinl(a);
        }
        t[k] = make_double2((double) c, (double) (-s));
    }
}

// [x2][k1] = w_nx^(k1*x2), nx = n1*n2, from the length-nx table
void make_interlevel(std::vector<double2> &w2, const std::vector<double2> &t, unsigned n1, unsigned n2)
{
    const size_t nx = (size_t) n1 * n2;
    w2.resize(nx);
    for (size_t x2 = 0; x2 < n2; ++x2)
        for (size_t k1 = 0; k1 < n1; ++k1) w2[x2 * n1 + k1] = t[(k1 * x2) % nx];
}

}  // namespace

struct hpxfft_b200_plan {
    int rank = 0, P = 1, device = 0, mode = MODE_SHARED;
    size_t nxl = 0, n_col = 0, ny = 0, cy = 0, nx = 0, m = 0;
    // column ownership
    unsigned wq0 = 0, w = 0, c0 = 0, ntiles = 0;
    std::vector<unsigned> ntiles_of, w_of, c0_of;
    // column FFT decomposition
    unsigned n1 = 1, n2 = 1;
    bool two_level = false;
    bool rows_generic = false, cols_generic = false; // direct-DFT kernels for lengths that are not powers of two
    // device buffers
    double *V = nullptr;   // slab, nxl x n_col doubles
    cd *bufA = nullptr;    // send buffer of exchange #1 and #2 (nranks > 1, NCCL modes)
    cd *bufB = nullptr;    // I (intermediate / receive window of exchange #1); receive buffer of #2
    cd *zraw = nullptr;    // un-split row spectra, only for rows longer than 32768 reals
    cd *S = nullptr;       // four-step scratch (full array, or an L2-resident ring of strips when fused)
    bool fused = false;    // level A + level B in one persistent launch
    bool fused_tma = false; // ... with the warp-specialised TMA-bulk / mbarrier pipeline
    unsigned lag = 0, nslot = 0, fused_grid = 0;
    unsigned *ctl = nullptr; // tile counter + per-strip completion counters
    cd *tw_row = nullptr, *tw_col = nullptr;
    cd *tw_il = nullptr;   // inter-level twiddles of the four-step column FFT, [x2][k1] = w_nx^(k1*x2)
    size_t bytesA = 0, bytesB = 0, bytesS = 0;
    // Experimental (single rank only): extra rows of padding between the column tiles of I, so that the tile
    // stride is not a power of two (suspected DRAM channel aliasing of the 256-byte segment traffic).
    size_t tile_pad_rows = 0;
    // p2p
    std::vector<void *> peerI, peerV;
    bool ipc_imported = false;
    // pipelined exchange (NCCL modes): sub-slab chunks of rows / strips, communication on its own stream
    int chunks_r = 1, chunks_c = 1;
    cd *bufC = nullptr;             // receive buffer of exchange #2 when it overlaps the column pass
    cudaStream_t cstream = nullptr; // high-priority communication stream
    std::vector<cudaEvent_t> ev_chunk; // [chunks_r + chunks_c + 2]
    int sm_reserve = 0;             // SMs left free for NCCL's kernels while the row kernel runs
    // execution
    cudaStream_t stream = nullptr;
    // One event set per execute since the last reset (ring of EV_SETS): lets a benchmark launch K
    // transforms back to back and still read per-phase / per-kernel averages afterwards.
    static constexpr int EV_SETS = 64, EV_PER_SET = 7;
    std::vector<cudaEvent_t> evs;   // EV_SETS * EV_PER_SET
    cudaEvent_t ev_io[2] = {};      // upload / download timing
    long nrec = 0;                  // executes enqueued since the last reset
    ncclComm_t comm = nullptr;
    int *d_barrier = nullptr;
    int launches = 0;
    std::map<std::string, double> meas;
    std::string plan_flag, row_desc, col_desc, col_desc_extra;
};

namespace {

// ------------------------------------------------------------------------------------------------
// kernel dispatch
// ------------------------------------------------------------------------------------------------
template <class K> int set_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024) CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
    return 0;
}

template <int M, int C, bool FAST> int launch_rows_big_t(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch)
{
    constexpr size_t smem = row_smem_total<M>();
    static int configured = -1;
    if (configured != p->device) {
        if (int rc = set_smem(rows_r2c_kernel<M, C, FAST>, smem)) return rc;
        configured = p->device;
    }
    const unsigned ngroups = (nrows + row_group<M>() - 1) / row_group<M>();
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
    // one resident CTA per SM (shared memory bound): persistent CTAs amortise the twiddle-table build
    sms -= p->sm_reserve;
    const unsigned cap = (unsigned) sms / C > 0 ? (unsigned) sms / C : 1u;
    const unsigned grid = ngroups < cap ? ngroups : cap;
    if (C > 2 && !p->zraw) return fail(HPXFFT_B200_ESTATE, "long-row scratch missing");
    rows_r2c_kernel<M, C, FAST><<<dim3(grid, C), ROW_THREADS, smem, p->stream>>>(V, pitch, nrows, dst, p->tw_row, p->zraw);
    CU(cudaGetLastError());
    if (C > 2) {
        const unsigned m = (unsigned) M * C;
        herm_split_kernel<<<dim3(nrows, (m / 2 + 1 + 255) / 256), 256, 0, p->stream>>>(p->zraw, m, nrows, dst, p->tw_row);
        CU(cudaGetLastError());
    }
    return 0;
}

template <int M, int C = 1> int launch_rows_big(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch)
{
    // fast output addressing: one destination rank and tile-aligned per-s stride (the 1-GPU hot configs)
    if constexpr (M == 8192 && C <= 2) {
        if (dst.P == 1) return launch_rows_big_t<M, C, true>(p, dst, nrows, V, pitch);
    }
    return launch_rows_big_t<M, C, false>(p, dst, nrows, V, pitch);
}

// EXPERIMENTAL 512-thread row kernel (kernels_rows16.cuh), opt-in with HPXFFT_B200_ROWS16=1
bool rows16_enabled()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("HPXFFT_B200_ROWS16");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

template <bool FAST> int launch_rows16_t(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch)
{
    static int configured = -1;
    if (configured != p->device) {
        if (int rc = set_smem(rows16_r2c_kernel<FAST>, rows16::SMEM)) return rc;
        configured = p->device;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
    sms -= p->sm_reserve;
    const unsigned grid = nrows < (unsigned) sms ? nrows : (unsigned) sms;
    rows16_r2c_kernel<FAST><<<grid, rows16::T, rows16::SMEM, p->stream>>>(V, pitch, nrows, dst, p->tw_row);
    CU(cudaGetLastError());
    return 0;
}

template <int M> int launch_rows_tiny(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch)
{
    const unsigned block = 128, grid = (nrows + block - 1) / block;
    rows_r2c_tiny_kernel<M><<<grid, block, 0, p->stream>>>(V, pitch, nrows, dst, p->tw_row);
    CU(cudaGetLastError());
    return 0;
}

int launch_rows(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch, size_t m)
{
    if (p->rows_generic) {
        const unsigned ny = (unsigned) (2 * m), cy = ny / 2 + 1;
        rows_generic_kernel<<<dim3(nrows, (cy + 127) / 128), 128, 0, p->stream>>>((const double *) V, 2 * pitch, nrows, ny, dst, p->tw_row);
        CU(cudaGetLastError());
        return 0;
    }
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
    case 8192:
        if (rows16_enabled()) return dst.P == 1 ? launch_rows16_t<true>(p, dst, nrows, V, pitch) : launch_rows16_t<false>(p, dst, nrows, V, pitch);
        return launch_rows_big<8192>(p, dst, nrows, V, pitch);
    case 16384: return launch_rows_big<8192, 2>(p, dst, nrows, V, pitch);
    case 32768: return launch_rows_big<8192, 4>(p, dst, nrows, V, pitch);
    case 65536: return launch_rows_big<8192, 8>(p, dst, nrows, V, pitch);
    default: return fail(HPXFFT_B200_EINVAL, "unsupported row length ny=%zu (ny/2 must be a power of two <= 65536)", 2 * m);
    }
}

template <int N> int launch_cols_single(const hpxfft_b200_plan *p, const InterView &in, const ColDst &out, unsigned ntiles)
{
    constexpr size_t smem = single_smem_bytes<N>();
    static int configured = -1;
    if (configured != p->device) {
        if (int rc = set_smem(cols_single_kernel<N>, smem)) return rc;
        configured = p->device;
    }
    cols_single_kernel<N><<<ntiles, col_threads(N), smem, p->stream>>>(in, out, p->tw_col);
    CU(cudaGetLastError());
    return 0;
}

template <int N1> int launch_cols_A(const hpxfft_b200_plan *p, const InterView &in, cd *S, unsigned n2, unsigned ntiles)
{
    constexpr size_t smem = levelA_smem_bytes<N1>();
    static int configured = -1;
    if (configured != p->device) {
        if (int rc = set_smem(cols_levelA_kernel<N1>, smem)) return rc;
        configured = p->device;
    }
    cols_levelA_kernel<N1><<<dim3(n2, ntiles), col_threads(N1), smem, p->stream>>>(in, S, n2, p->tw_col, p->tw_il);
    CU(cudaGetLastError());
    return 0;
}

template <int N2> int launch_cols_B(const hpxfft_b200_plan *p, const cd *S, const ColDst &out, unsigned n1, unsigned ntiles)
{
    constexpr size_t smem = single_smem_bytes<N2>();
    static int configured = -1;
    if (configured != p->device) {
        if (int rc = set_smem(cols_levelB_kernel<N2>, smem)) return rc;
        configured = p->device;
    }
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

int launch_cols(const hpxfft_b200_plan *p, const InterView &in, const ColDst &out, unsigned ntiles, cd *S, unsigned nx,
                unsigned n1, unsigned n2, bool two_level, int *launches, cudaEvent_t mid = nullptr)
{
    if (p->cols_generic) {
        if (launches) *launches += 1;
        if (mid) CU(cudaEventRecord(mid, p->stream));
        cols_generic_kernel<<<dim3(ntiles * CW, (nx + 127) / 128), 128, 0, p->stream>>>(in, out, nx, p->tw_col);
        CU(cudaGetLastError());
        return 0;
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

template <int N1, int N2> int fused_occupancy(int *blocks_per_sm, bool tma)
{
    if (tma) {
        constexpr size_t smem = tma_smem_bytes<N1, N2>();
        if (int rc = set_smem(cols_fused_tma_kernel<N1, N2>, smem)) return rc;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, cols_fused_tma_kernel<N1, N2>, tma_threads<N1, N2>(), smem));
        return 0;
    }
    constexpr size_t smem = fused_smem_bytes<N1, N2>();
    if (int rc = set_smem(cols_fused_kernel<N1, N2>, smem)) return rc;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, cols_fused_kernel<N1, N2>, fused_threads<N1, N2>(), smem));
    return 0;
}

template <int N1, int N2>
int launch_cols_fused_t(const hpxfft_b200_plan *p, const InterView &in, const ColDst &out, unsigned ct0, unsigned ntiles)
{
    constexpr size_t smem = fused_smem_bytes<N1, N2>();
    static int configured = -1;
    if (configured != p->device) {
        if (int rc = set_smem(cols_fused_kernel<N1, N2>, smem)) return rc;
        if (int rc = set_smem(cols_fused_tma_kernel<N1, N2>, tma_smem_bytes<N1, N2>())) return rc;
        configured = p->device;
    }
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
        static int discard = -1;
        if (discard < 0) {
            const char *e = getenv("HPXFFT_B200_DISCARD");
            discard = (e && e[0] == '0') ? 0 : 1;
        }
        ctl.discard = (unsigned) discard;
    }
    if (p->fused_tma)
        cols_fused_tma_kernel<N1, N2><<<p->fused_grid, tma_threads<N1, N2>(), tma_smem_bytes<N1, N2>(), p->stream>>>(in, p->S, out, p->tw_col,
                                                                                                          p->tw_il, ntiles, ctl);
    else
        cols_fused_kernel<N1, N2><<<p->fused_grid, fused_threads<N1, N2>(), smem, p->stream>>>(in, p->S, out, p->tw_col, p->tw_il, ntiles, ctl);
    CU(cudaGetLastError());
    return 0;
}

#define FUSED_PAIRS(X) X(32, 16) X(32, 32) X(64, 32) X(64, 64) X(128, 64) X(128, 128) X(256, 128) X(256, 256) X(512, 256) X(512, 512)

int launch_cols_fused(const hpxfft_b200_plan *p, const InterView &in, const ColDst &out, unsigned ct0, unsigned ntiles)
{
    if (ntiles == 0) return 0;
#define X(A, B) if (p->n1 == A && p->n2 == B) return launch_cols_fused_t<A, B>(p, in, out, ct0, ntiles);
    FUSED_PAIRS(X)
#undef X
    return fail(HPXFFT_B200_EINVAL, "no fused column kernel for %u x %u", p->n1, p->n2);
}

int fused_blocks_per_sm(unsigned n1, unsigned n2, int *bps, bool tma)
{
#define X(A, B) if (n1 == A && n2 == B) return fused_occupancy<A, B>(bps, tma);
    FUSED_PAIRS(X)
#undef X
    return fail(HPXFFT_B200_EINVAL, "no fused column kernel for %u x %u", n1, n2);
}

void choose_col_split(size_t nx, unsigned &n1, unsigned &n2, bool &two_level)
{
    if (nx <= 256) {
        two_level = false;
        n1 = (unsigned) nx;
        n2 = 1;
        return;
    }
    two_level = true;
    int lg = 0;
    while (((size_t) 1 << lg) < nx) ++lg;
    n1 = 1u << ((lg + 1) / 2);
    n2 = 1u << (lg / 2);
}

int parse_plan_flag(const char *f)
{
    if (!f) return -1;
    // core/include/hpxfft/util/adapter_fftw.hpp:22-44
    if (!strcmp(f, "estimate") || !strcmp(f, "measure") || !strcmp(f, "patient") || !strcmp(f, "exhaustive")) return 0;
    return -1;
}

int parse_comm_flag(const char *f, int *mode)
{
    if (!f) { *mode = MODE_SHARED; return 0; }
    if (!strcmp(f, "scatter")) { *mode = MODE_SCATTER; return 0; }
    if (!strcmp(f, "all_to_all")) { *mode = MODE_ALL_TO_ALL; return 0; }
    if (!strcmp(f, "p2p")) { *mode = MODE_P2P; return 0; }
    return -1;
}

int fill_rowdst(const hpxfft_b200_plan *p, RowDst &d)
{
    d.tile_stride = (unsigned long long) (p->nxl + p->tile_pad_rows) * CW;
    d.cy = (unsigned) p->cy;
    d.wq0 = p->wq0;
    d.P = (unsigned) p->P;
    unsigned long long off = 0;
    for (int q = 0; q < p->P; ++q) {
        const unsigned long long blk = (unsigned long long) p->ntiles_of[q] * p->nxl * CW;
        if (q == p->rank)
            d.base[q] = p->bufB + (unsigned long long) p->rank * ((unsigned long long) p->ntiles * p->nxl * CW);
        else if (p->mode == MODE_P2P)
            d.base[q] = (cd *) p->peerI[q] + (unsigned long long) p->rank * blk;
        else
            d.base[q] = p->bufA + off;
        off += blk;
    }
    return 0;
}

void fill_coldst(const hpxfft_b200_plan *p, ColDst &d)
{
    d.nxl = (unsigned) p->nxl;
    d.shift = pow2_shift(d.nxl);
    d.w = p->w;
    for (int r = 0; r < p->P; ++r) {
        if (r == p->rank) {
            d.base[r] = (cd *) p->V;
            d.pitch[r] = (unsigned) p->cy;
            d.col0[r] = (int) p->c0;
        } else if (p->mode == MODE_P2P) {
            d.base[r] = (cd *) p->peerV[r];
            d.pitch[r] = (unsigned) p->cy;
            d.col0[r] = (int) p->c0;
        } else {
            d.base[r] = p->bufA + (unsigned long long) r * p->nxl * p->w;
            d.pitch[r] = p->w;
            d.col0[r] = 0;
        }
    }
}

bool a2a_stepwise()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("HPXFFT_B200_A2A_STEPWISE");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

int barrier_on_stream(hpxfft_b200_plan *p)
{
    NC(g_nccl.AllReduce(p->d_barrier, p->d_barrier, 1, ncclInt, ncclSum, p->comm, p->stream));
    return 0;
}

// exchange #1: block (r -> q) = I-layout tiles of rank q's columns for my rows
int exchange1(hpxfft_b200_plan *p)
{
    const int P = p->P, me = p->rank;
    std::vector<unsigned long long> soff(P + 1, 0);
    for (int q = 0; q < P; ++q) soff[q + 1] = soff[q] + (unsigned long long) p->ntiles_of[q] * p->nxl * CW;
    const unsigned long long rblk = (unsigned long long) p->ntiles * p->nxl * CW; // what every peer sends me
    if (p->mode == MODE_ALL_TO_ALL) {
        // rotation schedule: step s pairs every rank with (me+s) / (me-s).  Either all P-1 steps in one NCCL
        // group (default) or one group per step (HPXFFT_B200_A2A_STEPWISE=1), which keeps each NVLink
        // transfer at full per-pair channel count when P is large.
        if (!a2a_stepwise()) NC(g_nccl.GroupStart());
        for (int s = 1; s < P; ++s) {
            const int to = (me + s) % P, from = (me - s + P) % P;
            if (a2a_stepwise()) NC(g_nccl.GroupStart());
            NC(g_nccl.Send(p->bufA + soff[to], (soff[to + 1] - soff[to]) * 2, ncclDouble, to, p->comm, p->stream));
            NC(g_nccl.Recv(p->bufB + (unsigned long long) from * rblk, rblk * 2, ncclDouble, from, p->comm, p->stream));
            if (a2a_stepwise()) NC(g_nccl.GroupEnd());
        }
        if (!a2a_stepwise()) NC(g_nccl.GroupEnd());
    } else { // scatter: one rooted scatter per locality, all in flight together like the reference's
             // asynchronous scatter_to / scatter_from futures (core/src/distributed/loop.cpp:158-167)
        NC(g_nccl.GroupStart());
        for (int root = 0; root < P; ++root) {
            if (root == me) {
                for (int to = 0; to < P; ++to)
                    if (to != me)
                        NC(g_nccl.Send(p->bufA + soff[to], (soff[to + 1] - soff[to]) * 2, ncclDouble, to, p->comm, p->stream));
            } else {
                NC(g_nccl.Recv(p->bufB + (unsigned long long) root * rblk, rblk * 2, ncclDouble, root, p->comm, p->stream));
            }
        }
        NC(g_nccl.GroupEnd());
    }
    return 0;
}

// exchange #2: block (q -> r) = dense [nxl][w_q] result rows of rank r
int exchange2(hpxfft_b200_plan *p)
{
    const int P = p->P, me = p->rank;
    const unsigned long long sblk = (unsigned long long) p->nxl * p->w;
    if (p->mode == MODE_ALL_TO_ALL) {
        if (!a2a_stepwise()) NC(g_nccl.GroupStart());
        for (int s = 1; s < P; ++s) {
            const int to = (me + s) % P, from = (me - s + P) % P;
            if (a2a_stepwise()) NC(g_nccl.GroupStart());
            NC(g_nccl.Send(p->bufA + (unsigned long long) to * sblk, sblk * 2, ncclDouble, to, p->comm, p->stream));
            NC(g_nccl.Recv(p->bufB + (unsigned long long) p->nxl * p->c0_of[from], (unsigned long long) p->nxl * p->w_of[from] * 2,
                        ncclDouble, from, p->comm, p->stream));
            if (a2a_stepwise()) NC(g_nccl.GroupEnd());
        }
        if (!a2a_stepwise()) NC(g_nccl.GroupEnd());
    } else {
        NC(g_nccl.GroupStart());
        for (int root = 0; root < P; ++root) {
            if (root == me) {
                for (int to = 0; to < P; ++to)
                    if (to != me)
                        NC(g_nccl.Send(p->bufA + (unsigned long long) to * sblk, sblk * 2, ncclDouble, to, p->comm, p->stream));
            } else {
                NC(g_nccl.Recv(p->bufB + (unsigned long long) p->nxl * p->c0_of[root],
                            (unsigned long long) p->nxl * p->w_of[root] * 2, ncclDouble, root, p->comm, p->stream));
            }
        }
        NC(g_nccl.GroupEnd());
    }
    return 0;
}

// strip range [t0, t1) and column range of chunk t when `ntiles` strips / `w` columns are cut into `nch` chunks
void chunk_bounds(unsigned ntiles, unsigned w, int nch, int t, unsigned &t0, unsigned &t1, unsigned &col0, unsigned &wc)
{
    const unsigned per = (ntiles + nch - 1) / nch;
    t0 = (unsigned) t * per < ntiles ? (unsigned) t * per : ntiles;
    t1 = (unsigned) (t + 1) * per < ntiles ? (unsigned) (t + 1) * per : ntiles;
    col0 = t0 * CW;
    const unsigned cend = t1 * CW < w ? t1 * CW : w;
    wc = cend > col0 ? cend - col0 : 0;
}

// group of sends/recvs of one exchange chunk; `order` = rotation (all_to_all) or root-major (scatter)
template <class SendFn, class RecvFn> int exchange_group(hpxfft_b200_plan *p, SendFn send, RecvFn recv)
{
    const int P = p->P, me = p->rank;
    NC(g_nccl.GroupStart());
    if (p->mode == MODE_ALL_TO_ALL) {
        for (int s = 1; s < P; ++s) {
            if (int rc = send((me + s) % P)) return rc;
            if (int rc = recv((me - s + P) % P)) return rc;
        }
    } else {
        for (int root = 0; root < P; ++root) {
            if (root == me) {
                for (int to = 0; to < P; ++to)
                    if (to != me)
                        if (int rc = send(to)) return rc;
            } else if (int rc = recv(root))
                return rc;
        }
    }
    NC(g_nccl.GroupEnd());
    return 0;
}

// NCCL modes with sub-slab pipelining: the exchange of row chunk s overlaps the row FFTs of chunk s+1,
// the exchange (+ unpack) of strip chunk t overlaps the column FFTs of chunk t+1.
int enqueue_transform_pipelined(hpxfft_b200_plan *p)
{
    const int P = p->P, me = p->rank, Sr = p->chunks_r, Sc = p->chunks_c;
    int launches = 0;
    const size_t nxs = p->nxl / Sr;
    cudaEvent_t *ev = p->evs.data() + (size_t) (p->nrec % hpxfft_b200_plan::EV_SETS) * hpxfft_b200_plan::EV_PER_SET;
    p->nrec += 1;
    cudaEvent_t *evr = p->ev_chunk.data(), *evc = p->ev_chunk.data() + Sr;
    cudaEvent_t ev_x1 = p->ev_chunk[Sr + Sc], ev_x2 = p->ev_chunk[Sr + Sc + 1];

    std::vector<unsigned long long> soff(P + 1, 0);
    for (int q = 0; q < P; ++q) soff[q + 1] = soff[q] + (unsigned long long) p->ntiles_of[q] * p->nxl * CW;
    const unsigned long long rblk = (unsigned long long) p->ntiles * p->nxl * CW;

    CU(cudaEventRecord(ev[0], p->stream));
    // the communication stream must not start before everything previously enqueued on the main stream
    CU(cudaStreamWaitEvent(p->cstream, ev[0], 0));
    // ---- dimension 1: row chunks, each followed by its exchange on the communication stream
    for (int s = 0; s < Sr; ++s) {
        RowDst rd;
        rd.tile_stride = (unsigned long long) nxs * CW;
        rd.cy = (unsigned) p->cy;
        rd.wq0 = p->wq0;
        rd.P = (unsigned) P;
        for (int q = 0; q < P; ++q) {
            const unsigned long long cs = (unsigned long long) p->ntiles_of[q] * nxs * CW; // chunk stride inside q's block
            rd.base[q] = (q == me ? p->bufB + (unsigned long long) me * rblk : p->bufA + soff[q]) + (unsigned long long) s * cs;
        }
        if (int rc = launch_rows(p, rd, (unsigned) nxs, (const cd *) p->V + (size_t) s * nxs * p->cy, (unsigned) p->cy, p->m)) return rc;
        launches += 1 + (p->m > 16384 ? 1 : 0);
        CU(cudaEventRecord(evr[s], p->stream));
        CU(cudaStreamWaitEvent(p->cstream, evr[s], 0));
        auto send = [&](int to) -> int {
            const unsigned long long cs = (unsigned long long) p->ntiles_of[to] * nxs * CW;
            NC(g_nccl.Send(p->bufA + soff[to] + (unsigned long long) s * cs, cs * 2, ncclDouble, to, p->comm, p->cstream));
            return 0;
        };
        auto recv = [&](int from) -> int {
            const unsigned long long cs = (unsigned long long) p->ntiles * nxs * CW;
            NC(g_nccl.Recv(p->bufB + (unsigned long long) from * rblk + (unsigned long long) s * cs, cs * 2, ncclDouble, from, p->comm,
                           p->cstream));
            return 0;
        };
        if (int rc = exchange_group(p, send, recv)) return rc;
    }
    CU(cudaEventRecord(ev[1], p->stream));
    CU(cudaEventRecord(ev_x1, p->cstream));
    CU(cudaStreamWaitEvent(p->stream, ev_x1, 0));
    CU(cudaEventRecord(ev[2], p->stream));
    CU(cudaEventRecord(ev[6], p->stream));

    // ---- dimension 2: strip chunks
    InterView iv;
    iv.base = p->bufB;
    iv.nxl = (unsigned) nxs; // I is chunk-major: [r][s][ct][js][c] == [x / nxs][ct][x % nxs][c]
    iv.shift = pow2_shift(iv.nxl);
    iv.tile_stride = (unsigned long long) nxs * CW;
    iv.rank_stride = (unsigned long long) p->ntiles * nxs * CW;
    for (int t = 0; t < Sc; ++t) {
        unsigned t0, t1, col0, wc;
        chunk_bounds(p->ntiles, p->w, Sc, t, t0, t1, col0, wc);
        if (t1 > t0) {
            ColDst cdst;
            cdst.nxl = (unsigned) p->nxl;
            cdst.shift = pow2_shift(cdst.nxl);
            cdst.w = p->w;
            const unsigned long long boff = (unsigned long long) P * p->nxl * col0;
            for (int r = 0; r < P; ++r) {
                if (r == me) {
                    cdst.base[r] = (cd *) p->V;
                    cdst.pitch[r] = (unsigned) p->cy;
                    cdst.col0[r] = (int) p->c0;
                } else {
                    cdst.base[r] = p->bufA + boff + (unsigned long long) r * p->nxl * wc;
                    cdst.pitch[r] = wc;
                    cdst.col0[r] = -(int) col0;
                }
            }
            if (Sc == 1 && !p->fused) {
                if (int rc = launch_cols(p, iv, cdst, p->ntiles, p->S, (unsigned) p->nx, p->n1, p->n2, p->two_level, &launches, nullptr)) return rc;
            } else {
                if (int rc = launch_cols_fused(p, iv, cdst, t0, t1 - t0)) return rc;
                launches += 1;
            }
        }
        CU(cudaEventRecord(evc[t], p->stream));
        CU(cudaStreamWaitEvent(p->cstream, evc[t], 0));
        UnpackChunk u;
        bool any = false;
        for (int q = 0; q < P; ++q) {
            unsigned q0, q1, qcol0, qwc;
            chunk_bounds(p->ntiles_of[q], p->w_of[q], Sc, t, q0, q1, qcol0, qwc);
            u.src_off[q] = (unsigned long long) p->nxl * (p->c0_of[q] + qcol0);
            u.dst_col[q] = p->c0_of[q] + qcol0;
            u.wc[q] = q == me ? 0 : qwc;
            any = any || u.wc[q] > 0;
        }
        auto send = [&](int to) -> int {
            if (wc == 0) return 0;
            const unsigned long long boff = (unsigned long long) P * p->nxl * col0;
            NC(g_nccl.Send(p->bufA + boff + (unsigned long long) to * p->nxl * wc, (unsigned long long) p->nxl * wc * 2, ncclDouble, to,
                           p->comm, p->cstream));
            return 0;
        };
        auto recv = [&](int from) -> int {
            if (u.wc[from] == 0) return 0;
            NC(g_nccl.Recv(p->bufC + u.src_off[from], (unsigned long long) p->nxl * u.wc[from] * 2, ncclDouble, from, p->comm, p->cstream));
            return 0;
        };
        if (int rc = exchange_group(p, send, recv)) return rc;
        if (any) {
            unpack_chunk_kernel<<<dim3((unsigned) p->nxl, (unsigned) P), 256, 0, p->cstream>>>(p->bufC, (cd *) p->V, (unsigned) p->cy, u);
            CU(cudaGetLastError());
            launches += 1;
        }
    }
    CU(cudaEventRecord(ev[3], p->stream));
    CU(cudaEventRecord(ev_x2, p->cstream));
    CU(cudaStreamWaitEvent(p->stream, ev_x2, 0));
    CU(cudaEventRecord(ev[4], p->stream));
    CU(cudaEventRecord(ev[5], p->stream));
    p->launches = launches;
    return 0;
}

int enqueue_transform(hpxfft_b200_plan *p)
{
    if (p->mode == MODE_P2P && p->P > 1 && !p->ipc_imported)
        return fail(HPXFFT_B200_ESTATE, "p2p plan: hpxfft_b200_ipc_import has not been called");
    if (p->P > 1 && p->mode != MODE_P2P && (p->chunks_r > 1 || p->chunks_c > 1)) return enqueue_transform_pipelined(p);
    int launches = 0;
    RowDst rd;
    fill_rowdst(p, rd);
    ColDst cdst;
    fill_coldst(p, cdst);
    InterView iv;
    iv.base = p->bufB;
    iv.nxl = (unsigned) p->nxl;
    iv.shift = pow2_shift(iv.nxl);
    iv.tile_stride = (unsigned long long) (p->nxl + p->tile_pad_rows) * CW;
    iv.rank_stride = (unsigned long long) p->ntiles * (p->nxl + p->tile_pad_rows) * CW;

    cudaEvent_t *ev = p->evs.data() + (size_t) (p->nrec % hpxfft_b200_plan::EV_SETS) * hpxfft_b200_plan::EV_PER_SET;
    p->nrec += 1;
    CU(cudaEventRecord(ev[0], p->stream));
    // phase 1: r2c rows (+ fused split / transpose)           -> first_fftw (first_split, first_trans fused)
    if (int rc = launch_rows(p, rd, (unsigned) p->nxl, (const cd *) p->V, (unsigned) p->cy, p->m)) return rc;
    launches += 1 + (p->m > 16384 ? 1 : 0);
    CU(cudaEventRecord(ev[1], p->stream));
    // phase 2: exchange #1                                    -> first_comm
    if (p->P > 1) {
        if (p->mode == MODE_P2P) {
            if (int rc = barrier_on_stream(p)) return rc;
        } else if (int rc = exchange1(p))
            return rc;
    }
    CU(cudaEventRecord(ev[2], p->stream));
    // phase 3: c2c columns (+ fused split / transpose)        -> second_fftw
    if (p->fused) {
        CU(cudaEventRecord(ev[6], p->stream));
        if (int rc = launch_cols_fused(p, iv, cdst, 0u, p->ntiles)) return rc;
        launches += 1;
    } else if (int rc = launch_cols(p, iv, cdst, p->ntiles, p->S, (unsigned) p->nx, p->n1, p->n2, p->two_level, &launches, ev[6]))
        return rc;
    CU(cudaEventRecord(ev[3], p->stream));
    // phase 4: exchange #2                                    -> second_comm
    if (p->P > 1) {
        if (p->mode == MODE_P2P) {
            if (int rc = barrier_on_stream(p)) return rc;
        } else if (int rc = exchange2(p))
            return rc;
    }
    CU(cudaEventRecord(ev[4], p->stream));
    // phase 5: unpack into the slab                           -> second_trans
    if (p->P > 1 && p->mode != MODE_P2P) {
        unpack_kernel<<<dim3((unsigned) p->nxl, (unsigned) p->P), 256, 0, p->stream>>>(p->bufB, (cd *) p->V, (unsigned) p->nxl,
                                                                                      (unsigned) p->cy, p->wq0, (unsigned) p->P, (unsigned) p->rank);
        CU(cudaGetLastError());
        launches += 1;
    }
    CU(cudaEventRecord(ev[5], p->stream));
    p->launches = launches;
    return 0;
}

int read_timers(hpxfft_b200_plan *p)
{
    // averages over the executes enqueued since the last reset (at most the EV_SETS most recent)
    const long nset = p->nrec < hpxfft_b200_plan::EV_SETS ? p->nrec : hpxfft_b200_plan::EV_SETS;
    if (nset <= 0) return 0;
    double acc[8] = {0};
    for (long sidx = 0; sidx < nset; ++sidx) {
        const long slot = ((p->nrec - 1 - sidx) % hpxfft_b200_plan::EV_SETS + hpxfft_b200_plan::EV_SETS) % hpxfft_b200_plan::EV_SETS;
        cudaEvent_t *ev = p->evs.data() + (size_t) slot * hpxfft_b200_plan::EV_PER_SET;
        float ms = 0;
        for (int i = 0; i < 5; ++i) {
            CU(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
            acc[i] += ms;
        }
        CU(cudaEventElapsedTime(&ms, ev[0], ev[5]));
        acc[5] += ms;
        CU(cudaEventElapsedTime(&ms, ev[2], ev[6]));
        acc[6] += ms; // column level A (0 for single-level)
        CU(cudaEventElapsedTime(&ms, ev[6], ev[3]));
        acc[7] += ms; // column level B / single-level kernel
    }
    const double sc = 1e-3 / (double) nset;
    auto &m = p->meas;
    m["total"] = acc[5] * sc;
    m["first_fftw"] = acc[0] * sc;
    m["first_split"] = 0.0; // fused into the row kernel's store
    m["first_comm"] = acc[1] * sc;
    m["first_trans"] = 0.0; // no transpose: the column kernel reads the tiled layout directly
    m["second_fftw"] = acc[2] * sc;
    m["second_split"] = 0.0;
    m["second_comm"] = acc[3] * sc;
    m["second_trans"] = acc[4] * sc;
    m["rows_kernel"] = acc[0] * sc;
    m["cols_kernel"] = acc[2] * sc;
    m["cols_levelA_kernel"] = acc[6] * sc;
    m["cols_levelB_kernel"] = acc[7] * sc;
    m["timer_samples"] = (double) nset;
    return 0;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int hpxfft_b200_version(void) { return HPXFFT_B200_VERSION; }
const char *hpxfft_b200_last_error(void) { return g_err; }

int hpxfft_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int hpxfft_b200_partition(size_t cy, int nranks, int rank, size_t *c0, size_t *w)
{
    if (nranks < 1 || rank < 0 || rank >= nranks || cy < (size_t) nranks || !c0 || !w)
        return fail(HPXFFT_B200_EINVAL, "bad partition request cy=%zu nranks=%d rank=%d", cy, nranks, rank);
    const size_t wq0 = cy / (size_t) nranks;
    *c0 = (size_t) rank * wq0;
    *w = (rank == nranks - 1) ? cy - *c0 : wq0;
    return 0;
}

int hpxfft_b200_get_unique_id(void *id_out)
{
    static_assert(sizeof(ncclUniqueId) == HPXFFT_B200_UNIQUE_ID_BYTES, "unique id size");
    if (!id_out) return fail(HPXFFT_B200_EINVAL, "id_out is NULL");
    if (int rc = nccl_load()) return rc;
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

void hpxfft_b200_destroy(hpxfft_b200_plan *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->ipc_imported) {
        for (int q = 0; q < p->P; ++q) {
            if (q == p->rank) continue;
            if (p->peerI[q]) cudaIpcCloseMemHandle(p->peerI[q]);
            if (p->peerV[q]) cudaIpcCloseMemHandle(p->peerV[q]);
        }
    }
    if (p->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(p->comm);
    for (auto &e : p->ev_chunk)
        if (e) cudaEventDestroy(e);
    if (p->cstream) {
        cudaStreamSynchronize(p->cstream);
        cudaStreamDestroy(p->cstream);
    }
    cudaFree(p->bufC);
    cudaFree(p->V);
    cudaFree(p->bufA);
    cudaFree(p->bufB);
    cudaFree(p->S);
    cudaFree(p->zraw);
    cudaFree(p->ctl);
    cudaFree(p->tw_row);
    cudaFree(p->tw_col);
    cudaFree(p->tw_il);
    cudaFree(p->d_barrier);
    for (auto &e : p->evs)
        if (e) cudaEventDestroy(e);
    for (auto &e : p->ev_io)
        if (e) cudaEventDestroy(e);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

int hpxfft_b200_create(hpxfft_b200_plan **out, size_t n_x_local, size_t n_col, int rank, int nranks, int device,
                       const char *comm_flag, const char *plan_flag, const void *unique_id)
{
    if (!out) return fail(HPXFFT_B200_EINVAL, "out is NULL");
    *out = nullptr;
    if (parse_plan_flag(plan_flag)) return fail(HPXFFT_B200_EPLANFLAG, "Invalid FFTW plan flag string");
    int mode = 0;
    if (parse_comm_flag(comm_flag, &mode))
        return fail(HPXFFT_B200_ECOMMFLAG, "Specify communication scheme: scatter or all_to_all");
    if (nranks < 1 || nranks > MAXP || rank < 0 || rank >= nranks) return fail(HPXFFT_B200_EINVAL, "bad rank/nranks %d/%d", rank, nranks);
    if (mode == MODE_SHARED && nranks != 1) return fail(HPXFFT_B200_EINVAL, "shared::loop needs exactly one locality");
    if (n_x_local == 0 || n_col < 4 || (n_col & 1)) return fail(HPXFFT_B200_EINVAL, "bad slab shape %zu x %zu", n_x_local, n_col);
    if (nranks > 1 && !unique_id) return fail(HPXFFT_B200_EINVAL, "unique_id required for nranks > 1");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(HPXFFT_B200_ECUDA, "no CUDA device available (libhpxfft_b200 has no CPU fallback)");
    }
    if (device < 0) CU(cudaGetDevice(&device));
    if (device >= ndev) return fail(HPXFFT_B200_EINVAL, "device %d out of range (%d devices)", device, ndev);
    CU(cudaSetDevice(device));

    hpxfft_b200_plan *p = new hpxfft_b200_plan();
    p->rank = rank;
    p->P = nranks;
    p->device = device;
    p->mode = mode;
    p->plan_flag = plan_flag;
    // dimension inference: core/src/shared/loop.cpp:163-165, core/src/distributed/loop.cpp:284-287
    p->nxl = n_x_local;
    p->n_col = n_col;
    p->cy = n_col / 2;
    p->ny = 2 * p->cy - 2;
    p->nx = n_x_local * (size_t) nranks;
    p->m = p->ny / 2;

    auto bail = [&](int rc) {
        hpxfft_b200_destroy(p);
        return rc;
    };
    // powers of two take the Stockham kernels; any other length (FFTW accepts them all, the reference's
    // default example is 8 x 14) takes the direct-DFT kernels, bounded to sizes where O(n^2) is sane
    constexpr size_t GENERIC_MAX = 8192;
    p->rows_generic = !is_pow2(p->m);
    p->cols_generic = !is_pow2(p->nx);
    if (p->ny < 2) return bail(fail(HPXFFT_B200_EINVAL, "unsupported ny=%zu", p->ny));
    if (p->rows_generic ? p->ny > GENERIC_MAX : p->m > 65536)
        return bail(fail(HPXFFT_B200_EINVAL, "unsupported ny=%zu: ny/2 must be a power of two <= 65536, or ny <= %zu", p->ny, GENERIC_MAX));
    if (p->cols_generic ? p->nx > GENERIC_MAX : p->nx > (1u << 18))
        return bail(fail(HPXFFT_B200_EINVAL, "unsupported nx=%zu: must be a power of two <= 2^18, or <= %zu", p->nx, GENERIC_MAX));
    if (p->cy < (size_t) nranks) return bail(fail(HPXFFT_B200_EINVAL, "ny/2+1=%zu columns cannot be split over %d localities", p->cy, nranks));

    // column ownership: c_q = q*floor(cy/P), the last rank absorbs cy mod P
    p->wq0 = (unsigned) (p->cy / nranks);
    p->ntiles_of.resize(nranks);
    p->w_of.resize(nranks);
    p->c0_of.resize(nranks);
    for (int q = 0; q < nranks; ++q) {
        p->c0_of[q] = q * p->wq0;
        p->w_of[q] = (q == nranks - 1) ? (unsigned) p->cy - p->c0_of[q] : p->wq0;
        p->ntiles_of[q] = (p->w_of[q] + CW - 1) / CW;
    }
    p->w = p->w_of[rank];
    p->c0 = p->c0_of[rank];
    p->ntiles = p->ntiles_of[rank];
    choose_col_split(p->nx, p->n1, p->n2, p->two_level);
    if (p->cols_generic) {
        p->two_level = false;
        p->n1 = (unsigned) p->nx;
        p->n2 = 1;
    }

    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (cudaEventCreate(&t0) != cudaSuccess || cudaEventCreate(&t1) != cudaSuccess) return bail(fail(HPXFFT_B200_ECUDA, "event create failed"));
    if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(HPXFFT_B200_ECUDA, "stream create failed"));
    p->evs.assign((size_t) hpxfft_b200_plan::EV_SETS * hpxfft_b200_plan::EV_PER_SET, nullptr);
    for (auto &e : p->evs)
        if (cudaEventCreate(&e) != cudaSuccess) return bail(fail(HPXFFT_B200_ECUDA, "event create failed"));
    for (auto &e : p->ev_io)
        if (cudaEventCreate(&e) != cudaSuccess) return bail(fail(HPXFFT_B200_ECUDA, "event create failed"));
    cudaEventRecord(t0, p->stream);

    // buffers
    const size_t bytesV = p->nxl * p->n_col * sizeof(double);
    if (nranks == 1) {
        if (const char *e = getenv("HPXFFT_B200_TILE_PAD")) { int v = atoi(e); if (v > 0 && v <= 4096) p->tile_pad_rows = (size_t) v; }
    }
    p->bytesB = (size_t) p->ntiles * (p->nx + p->tile_pad_rows) * CW * sizeof(cd); // I: [r][ct][j][c], also >= nxl*cy for exchange #2
    if (p->bytesB < p->nxl * p->cy * sizeof(cd)) p->bytesB = p->nxl * p->cy * sizeof(cd);
    p->bytesS = p->two_level ? (size_t) p->ntiles * p->nx * CW * sizeof(cd) : 0;
    if (p->two_level) {
        const char *e = getenv("HPXFFT_B200_FUSED");
        p->fused = !(e && e[0] == '0');
        p->fused_tma = p->fused && (e && e[0] == '2'); // HPXFFT_B200_FUSED=2 selects the TMA-bulk/mbarrier variant
        // N = 512 tiles need 2 x 128 KB with a staging buffer: fall back to the plain fused kernel
        if (p->n2 < 32) p->fused_tma = false; // single-pass tiles have no shared-memory buffer to refill
    }
    if (nranks > 1 && mode != MODE_P2P) {
        // Sub-slab pipelining of the NCCL exchanges is implemented and parity-tested but OFF by default:
        // NCCL's copy kernels need >= 32 CTAs for full NVLink rate, which the persistent FFT kernels
        // cannot spare without losing more than the overlap wins.
        int want = 1;
        if (const char *e = getenv("HPXFFT_B200_CHUNKS")) want = atoi(e) > 0 ? atoi(e) : 1;
        if (want > 1) {
            p->sm_reserve = 16;
            if (const char *e = getenv("HPXFFT_B200_SM_RESERVE")) { int v = atoi(e); if (v >= 0 && v <= 64) p->sm_reserve = v; }
        }
    }
    if (p->fused) {
        int bps = 1, sms = 148;
        if (int rc = fused_blocks_per_sm(p->n1, p->n2, &bps, p->fused_tma)) return bail(rc);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (const char *e = getenv("HPXFFT_B200_FUSED_BPS")) { int v = atoi(e); if (v >= 1 && v < bps) bps = v; }
        p->fused_grid = (unsigned) (bps * (sms - p->sm_reserve));
        const unsigned per_group = p->n1 + p->n2;
        p->lag = (unsigned) ((3 * (size_t) p->fused_grid / 2 + per_group - 1) / per_group) + 1;
        if (const char *e = getenv("HPXFFT_B200_LAG")) { int v = atoi(e); if (v >= 1) p->lag = (unsigned) v; }
        p->nslot = 2 * p->lag + 1;
        if (const char *e = getenv("HPXFFT_B200_NSLOT")) { int v = atoi(e); if (v > (int) p->lag) p->nslot = (unsigned) v; }
        if (p->nslot > p->ntiles) p->nslot = p->ntiles > 0 ? p->ntiles : 1;
        p->bytesS = (size_t) p->nslot * p->nx * CW * sizeof(cd);
    }
    if (nranks > 1 && mode != MODE_P2P) {
        size_t tiles_all = 0;
        for (int q = 0; q < nranks; ++q) tiles_all += p->ntiles_of[q];
        p->bytesA = tiles_all * p->nxl * CW * sizeof(cd);
        const size_t ex2 = (size_t) nranks * p->nxl * p->w * sizeof(cd);
        if (p->bytesA < ex2) p->bytesA = ex2;
    }
#define CUB(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return bail(fail(HPXFFT_B200_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)));     \
    } while (0)
    CUB(cudaMalloc(&p->V, bytesV));
    CUB(cudaMalloc(&p->bufB, p->bytesB));
    if (p->m > 16384) CUB(cudaMalloc(&p->zraw, p->nxl * p->m * sizeof(cd)));
    if (p->bytesS) CUB(cudaMalloc(&p->S, p->bytesS));
    if (p->fused) CUB(cudaMalloc(&p->ctl, (1 + 2 * (size_t) p->ntiles) * sizeof(unsigned)));
    if (p->bytesA) CUB(cudaMalloc(&p->bufA, p->bytesA));
    CUB(cudaMemsetAsync(p->V, 0, bytesV, p->stream));
    CUB(cudaMemsetAsync(p->bufB, 0, p->bytesB, p->stream));

    // "plans": twiddle tables w_ny^k and w_nx^k
    {
        std::vector<double2> t;
        make_twiddles(t, p->ny);
        CUB(cudaMalloc(&p->tw_row, t.size() * sizeof(double2)));
        CUB(cudaMemcpyAsync(p->tw_row, t.data(), t.size() * sizeof(double2), cudaMemcpyHostToDevice, p->stream));
        CUB(cudaStreamSynchronize(p->stream));
        make_twiddles(t, p->nx);
        CUB(cudaMalloc(&p->tw_col, t.size() * sizeof(double2)));
        CUB(cudaMemcpyAsync(p->tw_col, t.data(), t.size() * sizeof(double2), cudaMemcpyHostToDevice, p->stream));
        CUB(cudaStreamSynchronize(p->stream));
        if (p->two_level) {
            std::vector<double2> w2;
            make_interlevel(w2, t, p->n1, p->n2);
            CUB(cudaMalloc(&p->tw_il, w2.size() * sizeof(double2)));
            CUB(cudaMemcpyAsync(p->tw_il, w2.data(), w2.size() * sizeof(double2), cudaMemcpyHostToDevice, p->stream));
            CUB(cudaStreamSynchronize(p->stream));
        }
    }

    if (nranks > 1 && mode != MODE_P2P) {
        int want = 1;
        if (const char *e = getenv("HPXFFT_B200_CHUNKS")) want = atoi(e) > 0 ? atoi(e) : 1;
        int sr = want, sc = want;
        while (sr > 1 && (p->nxl % sr != 0 || p->nxl / sr < 8)) sr /= 2;
        if (!p->fused) sc = 1;
        while (sc > 1 && p->ntiles / sc < 2 * (p->lag + 1)) sc /= 2;
        p->chunks_r = sr < 1 ? 1 : sr;
        p->chunks_c = sc < 1 ? 1 : sc;
        if (p->chunks_r > 1 || p->chunks_c > 1) {
            CUB(cudaMalloc(&p->bufC, p->nxl * p->cy * sizeof(cd)));
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            CUB(cudaStreamCreateWithPriority(&p->cstream, cudaStreamNonBlocking, hi));
            p->ev_chunk.assign((size_t) p->chunks_r + p->chunks_c + 2, nullptr);
            for (auto &e : p->ev_chunk) CUB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
    }
    if (nranks > 1) {
        if (int rc = nccl_load()) return bail(rc);
        ncclUniqueId id;
        memcpy(&id, unique_id, sizeof(id));
        ncclResult_t r;
        if (p->sm_reserve > 0 && g_nccl.CommInitRankConfig) {
            // keep NCCL's kernels inside the SMs the FFT kernels leave free, so that the overlapped
            // exchange never evicts a persistent FFT CTA
            ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
            cfg.maxCTAs = p->sm_reserve;
            r = g_nccl.CommInitRankConfig(&p->comm, nranks, id, rank, &cfg);
        } else
            r = g_nccl.CommInitRank(&p->comm, nranks, id, rank);
        if (r != ncclSuccess) return bail(fail(HPXFFT_B200_ENCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)));
        CUB(cudaMalloc(&p->d_barrier, sizeof(int)));
        CUB(cudaMemsetAsync(p->d_barrier, 0, sizeof(int), p->stream));
        p->peerI.assign(nranks, nullptr);
        p->peerV.assign(nranks, nullptr);
    }

    cudaEventRecord(t1, p->stream);
    CUB(cudaStreamSynchronize(p->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, t1);
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    p->meas["plan"] = ms * 1e-3;
    // nominal flop count of the transform (the reference reports fftw_flops-based counts, shared/loop.cpp:188)
    const double N = (double) p->nx * (double) p->ny;
    p->meas["plan_flops"] = 2.5 * N * std::log2(N > 1 ? N : 2);

    char buf[256];
    snprintf(buf, sizeof(buf), "r2c rows: n=%zu via half-length complex Stockham m=%zu (%s), %d points/thread, paired radix-16 last pass + Hermitian split",
             p->ny, p->m, p->m <= 16 ? "register-resident" : "shared-memory pencil", p->m <= 16 ? (int) p->m : ROW_PT);
    p->row_desc = buf;
    if (p->rows_generic) p->row_desc = "r2c rows: direct DFT (length is not a power of two)";
    if (p->cols_generic)
        snprintf(buf, sizeof(buf), "c2c columns: n=%zu direct DFT (length is not a power of two)", p->nx);
    else if (p->two_level)
        snprintf(buf, sizeof(buf), "c2c columns: n=%zu four-step %u x %u on %d-column tiles (level A strided + twiddle, level B contiguous)%s",
                 p->nx, p->n1, p->n2, CW, p->fused ? ", fused persistent launch with L2-resident scratch ring" : "");
    if (p->fused_tma) p->col_desc_extra = " [producer warp: cp.async.bulk + mbarrier]";
    else
        snprintf(buf, sizeof(buf), "c2c columns: n=%zu single Stockham tile FFT on %d-column tiles", p->nx, CW);
    p->col_desc = buf;
    *out = p;
    return 0;
}

int hpxfft_b200_ipc_count(const hpxfft_b200_plan *p) { return (p && p->mode == MODE_P2P) ? 2 : 0; }

int hpxfft_b200_ipc_export(hpxfft_b200_plan *p, void *handles_out)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == HPXFFT_B200_IPC_HANDLE_BYTES, "ipc handle size");
    if (!p || !handles_out) return fail(HPXFFT_B200_EINVAL, "NULL argument");
    if (p->mode != MODE_P2P) return fail(HPXFFT_B200_ESTATE, "not a p2p plan");
    CU(cudaSetDevice(p->device));
    cudaIpcMemHandle_t h[2];
    CU(cudaIpcGetMemHandle(&h[0], p->bufB));
    CU(cudaIpcGetMemHandle(&h[1], p->V));
    memcpy(handles_out, h, sizeof(h));
    return 0;
}

int hpxfft_b200_ipc_import(hpxfft_b200_plan *p, const void *all_handles)
{
    if (!p || !all_handles) return fail(HPXFFT_B200_EINVAL, "NULL argument");
    if (p->mode != MODE_P2P) return fail(HPXFFT_B200_ESTATE, "not a p2p plan");
    CU(cudaSetDevice(p->device));
    const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *) all_handles;
    for (int q = 0; q < p->P; ++q) {
        if (q == p->rank) {
            p->peerI[q] = p->bufB;
            p->peerV[q] = p->V;
            continue;
        }
        CU(cudaIpcOpenMemHandle(&p->peerI[q], h[2 * q + 0], cudaIpcMemLazyEnablePeerAccess));
        CU(cudaIpcOpenMemHandle(&p->peerV[q], h[2 * q + 1], cudaIpcMemLazyEnablePeerAccess));
    }
    p->ipc_imported = true;
    return 0;
}

int hpxfft_b200_upload(hpxfft_b200_plan *p, const double *host_slab)
{
    if (!p || !host_slab) return fail(HPXFFT_B200_EINVAL, "NULL argument");
    CU(cudaSetDevice(p->device));
    cudaEvent_t a = p->ev_io[0], b = p->ev_io[1];
    CU(cudaEventRecord(a, p->stream));
    CU(cudaMemcpyAsync(p->V, host_slab, p->nxl * p->n_col * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CU(cudaEventRecord(b, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, a, b));
    p->meas["h2d"] = ms * 1e-3;
    return 0;
}

int hpxfft_b200_download(hpxfft_b200_plan *p, double *host_slab)
{
    if (!p || !host_slab) return fail(HPXFFT_B200_EINVAL, "NULL argument");
    CU(cudaSetDevice(p->device));
    cudaEvent_t a = p->ev_io[0], b = p->ev_io[1];
    CU(cudaEventRecord(a, p->stream));
    CU(cudaMemcpyAsync(host_slab, p->V, p->nxl * p->n_col * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaEventRecord(b, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, a, b));
    p->meas["d2h"] = ms * 1e-3;
    return 0;
}

int hpxfft_b200_fill(hpxfft_b200_plan *p, int pattern, uint64_t seed)
{
    if (!p) return fail(HPXFFT_B200_EINVAL, "NULL plan");
    if (pattern < 0 || pattern > 2) return fail(HPXFFT_B200_EINVAL, "unknown pattern %d", pattern);
    CU(cudaSetDevice(p->device));
    fill_kernel<<<148 * 8, 256, 0, p->stream>>>(p->V, (unsigned) p->nxl, (unsigned) p->ny, (unsigned) p->n_col,
                                               (unsigned long long) p->rank * p->nxl, pattern, seed);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(p->stream));
    return 0;
}

int hpxfft_b200_execute_async(hpxfft_b200_plan *p)
{
    if (!p) return fail(HPXFFT_B200_EINVAL, "NULL plan");
    CU(cudaSetDevice(p->device));
    return enqueue_transform(p);
}

int hpxfft_b200_synchronize(hpxfft_b200_plan *p)
{
    if (!p) return fail(HPXFFT_B200_EINVAL, "NULL plan");
    CU(cudaSetDevice(p->device));
    CU(cudaStreamSynchronize(p->stream));
    return read_timers(p);
}

int hpxfft_b200_reset_timers(hpxfft_b200_plan *p)
{
    if (!p) return fail(HPXFFT_B200_EINVAL, "NULL plan");
    CU(cudaSetDevice(p->device));
    CU(cudaStreamSynchronize(p->stream));
    p->nrec = 0;
    return 0;
}

int hpxfft_b200_execute(hpxfft_b200_plan *p)
{
    if (!p) return fail(HPXFFT_B200_EINVAL, "NULL plan");
    p->nrec = 0;
    if (int rc = hpxfft_b200_execute_async(p)) return rc;
    CU(cudaStreamSynchronize(p->stream));
    return read_timers(p);
}

int hpxfft_b200_transform(hpxfft_b200_plan *p, double *host_slab_inout)
{
    if (!p || !host_slab_inout) return fail(HPXFFT_B200_EINVAL, "NULL argument");
    CU(cudaSetDevice(p->device));
    const size_t bytes = p->nxl * p->n_col * sizeof(double);
    p->nrec = 0;
    CU(cudaMemcpyAsync(p->V, host_slab_inout, bytes, cudaMemcpyHostToDevice, p->stream));
    if (int rc = enqueue_transform(p)) return rc;
    CU(cudaMemcpyAsync(host_slab_inout, p->V, bytes, cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    return read_timers(p);
}

int hpxfft_b200_transform_async(hpxfft_b200_plan *p, double *host_slab_inout)
{
    if (!p || !host_slab_inout) return fail(HPXFFT_B200_EINVAL, "NULL argument");
    CU(cudaSetDevice(p->device));
    const size_t bytes = p->nxl * p->n_col * sizeof(double);
    CU(cudaMemcpyAsync(p->V, host_slab_inout, bytes, cudaMemcpyHostToDevice, p->stream));
    if (int rc = enqueue_transform(p)) return rc;
    CU(cudaMemcpyAsync(host_slab_inout, p->V, bytes, cudaMemcpyDeviceToHost, p->stream));
    return 0;
}

double hpxfft_b200_measurement(const hpxfft_b200_plan *p, const char *key)
{
    if (!p || !key) return 0.0;
    auto it = p->meas.find(key);
    return it == p->meas.end() ? 0.0 : it->second; // unknown key -> 0.0 (core/src/shared/loop.cpp:192)
}

int hpxfft_b200_write_plans(const hpxfft_b200_plan *p, const char *file_path)
{
    if (!p || !file_path) return fail(HPXFFT_B200_EINVAL, "NULL argument");
    FILE *f = fopen(file_path, "a");
    if (!f) return fail(HPXFFT_B200_EINVAL, "Failed to open file: %s", file_path);
    // same two-section structure as core/src/shared/loop.cpp:203-209
    fprintf(f, "FFTW r2c 1D plan:\n(hpxfft_b200 sm_100a %s)\n", p->row_desc.c_str());
    fprintf(f, "FFTW c2c 1D plan:\n(hpxfft_b200 sm_100a %s%s)\n\n", p->col_desc.c_str(), p->col_desc_extra.c_str());
    fclose(f);
    return 0;
}

void *hpxfft_b200_device_ptr(hpxfft_b200_plan *p) { return p ? p->V : nullptr; }
void *hpxfft_b200_stream(hpxfft_b200_plan *p) { return p ? (void *) p->stream : nullptr; }
int hpxfft_b200_launches_per_execute(const hpxfft_b200_plan *p)
{
    if (!p) return 0;
    if (p->launches > 0) return p->launches; // counted by the last execute
    int n = 1 + ((p->two_level && !p->fused) ? 2 : 1) + (p->m > 16384 ? 1 : 0);
    if (p->P > 1 && p->mode != MODE_P2P) n += 1;
    return n;
}

void *hpxfft_b200_host_alloc(size_t bytes)
{
    void *ptr = nullptr;
    cudaError_t e = cudaHostAlloc(&ptr, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(HPXFFT_B200_ECUDA, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return ptr;
}

void hpxfft_b200_host_free(void *ptr)
{
    if (ptr) cudaFreeHost(ptr);
}

// ---- adapter-level entry points (kernel parity tests) -------------------------------------------
int hpxfft_b200_r2c_rows(double *host_rows, size_t batch, size_t n_col, int device)
{
    if (!host_rows || batch == 0 || n_col < 4 || (n_col & 1)) return fail(HPXFFT_B200_EINVAL, "bad arguments");
    const size_t cy = n_col / 2, ny = 2 * cy - 2, m = ny / 2;
    if (!is_pow2(m) || m > 65536) return fail(HPXFFT_B200_EINVAL, "unsupported ny=%zu", ny);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(HPXFFT_B200_ECUDA, "no CUDA device available (libhpxfft_b200 has no CPU fallback)");
    }
    if (device < 0) CU(cudaGetDevice(&device));
    CU(cudaSetDevice(device));
    hpxfft_b200_plan P;
    hpxfft_b200_plan *p = &P;
    p->device = device;
    p->nxl = batch;
    p->cy = cy;
    p->ntiles = (unsigned) ((cy + CW - 1) / CW);
    const size_t bytes = batch * n_col * sizeof(double), tbytes = (size_t) p->ntiles * batch * CW * sizeof(cd);
    std::vector<double2> t;
    make_twiddles(t, ny);
    int rc = 0;
    auto cleanup = [&]() {
        cudaFree(p->V);
        cudaFree(p->bufB);
        cudaFree(p->tw_row);
        cudaFree(p->zraw);
        if (p->stream) cudaStreamDestroy(p->stream);
        p->V = nullptr; p->bufB = nullptr; p->tw_row = nullptr; p->zraw = nullptr; p->stream = nullptr;
    };
#define CUR(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            rc = fail(HPXFFT_B200_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_));                \
            cleanup();                                                                                   \
            return rc;                                                                                   \
        }                                                                                                \
    } while (0)
    CUR(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    CUR(cudaMalloc(&p->V, bytes));
    CUR(cudaMalloc(&p->bufB, tbytes));
    if (m > 16384) CUR(cudaMalloc(&p->zraw, batch * m * sizeof(cd)));
    CUR(cudaMalloc(&p->tw_row, t.size() * sizeof(double2)));
    CUR(cudaMemcpyAsync(p->tw_row, t.data(), t.size() * sizeof(double2), cudaMemcpyHostToDevice, p->stream));
    CUR(cudaMemcpyAsync(p->V, host_rows, bytes, cudaMemcpyHostToDevice, p->stream));
    RowDst rd;
    rd.tile_stride = (unsigned long long) batch * CW;
    rd.cy = (unsigned) cy;
    rd.wq0 = (unsigned) cy;
    rd.P = 1;
    rd.base[0] = p->bufB;
    rc = launch_rows(p, rd, (unsigned) batch, (const cd *) p->V, (unsigned) cy, m);
    if (rc) {
        cleanup();
        return rc;
    }
    untile_kernel<<<148 * 4, 256, 0, p->stream>>>(p->bufB, (cd *) p->V, (unsigned) batch, (unsigned) cy);
    CUR(cudaGetLastError());
    CUR(cudaMemcpyAsync(host_rows, p->V, bytes, cudaMemcpyDeviceToHost, p->stream));
    CUR(cudaStreamSynchronize(p->stream));
    cleanup();
    return 0;
}

int hpxfft_b200_c2c_cols(double *host_data, size_t n, size_t width, int device)
{
    if (!host_data || n == 0 || width == 0) return fail(HPXFFT_B200_EINVAL, "bad arguments");
    // a plan with nx = n rows and cy = width complex columns (n_col = 2*width); ny is irrelevant here,
    // so build the pieces by hand instead of going through create()'s ny checks
    if (!is_pow2(n) || n > (1u << 18)) return fail(HPXFFT_B200_EINVAL, "unsupported n=%zu", n);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(HPXFFT_B200_ECUDA, "no CUDA device available (libhpxfft_b200 has no CPU fallback)");
    }
    if (device < 0) CU(cudaGetDevice(&device));
    CU(cudaSetDevice(device));
    hpxfft_b200_plan P;
    hpxfft_b200_plan *p = &P;
    p->device = device;
    p->nxl = p->nx = n;
    p->cy = width;
    p->w = (unsigned) width;
    p->ntiles = (unsigned) ((width + CW - 1) / CW);
    choose_col_split(n, p->n1, p->n2, p->two_level);
    const size_t bytes = n * width * sizeof(cd), tbytes = (size_t) p->ntiles * n * CW * sizeof(cd);
    cd *A = nullptr;
    std::vector<double2> t;
    make_twiddles(t, n);
    int rc = 0;
    auto cleanup = [&]() {
        cudaFree(A);
        cudaFree(p->bufB);
        cudaFree(p->S);
        cudaFree(p->tw_col);
        cudaFree(p->tw_il);
        if (p->stream) cudaStreamDestroy(p->stream);
        p->bufB = nullptr; p->S = nullptr; p->tw_col = nullptr; p->tw_il = nullptr; p->stream = nullptr;
    };
#define CUC(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            rc = fail(HPXFFT_B200_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_));                \
            cleanup();                                                                                   \
            return rc;                                                                                   \
        }                                                                                                \
    } while (0)
    CUC(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    CUC(cudaMalloc(&A, bytes));
    CUC(cudaMalloc(&p->bufB, tbytes));
    if (p->two_level) CUC(cudaMalloc(&p->S, tbytes));
    CUC(cudaMalloc(&p->tw_col, t.size() * sizeof(double2)));
    CUC(cudaMemcpyAsync(p->tw_col, t.data(), t.size() * sizeof(double2), cudaMemcpyHostToDevice, p->stream));
    std::vector<double2> w2;
    if (p->two_level) {
        make_interlevel(w2, t, p->n1, p->n2);
        CUC(cudaMalloc(&p->tw_il, w2.size() * sizeof(double2)));
        CUC(cudaMemcpyAsync(p->tw_il, w2.data(), w2.size() * sizeof(double2), cudaMemcpyHostToDevice, p->stream));
    }
    CUC(cudaMemcpyAsync(A, host_data, bytes, cudaMemcpyHostToDevice, p->stream));
    tile_kernel<<<148 * 4, 256, 0, p->stream>>>(A, p->bufB, (unsigned) n, (unsigned) width);
    CUC(cudaGetLastError());
    InterView iv;
    iv.base = p->bufB;
    iv.nxl = (unsigned) n;
    iv.shift = pow2_shift(iv.nxl);
    iv.tile_stride = (unsigned long long) n * CW;
    iv.rank_stride = 0;
    ColDst cdst;
    cdst.nxl = (unsigned) n;
    cdst.shift = pow2_shift(cdst.nxl);
    cdst.w = (unsigned) width;
    cdst.base[0] = A;
    cdst.pitch[0] = (unsigned) width;
    cdst.col0[0] = 0;
    rc = launch_cols(p, iv, cdst, p->ntiles, p->S, (unsigned) n, p->n1, p->n2, p->two_level, nullptr);
    if (rc) {
        cleanup();
        return rc;
    }
    CUC(cudaMemcpyAsync(host_data, A, bytes, cudaMemcpyDeviceToHost, p->stream));
    CUC(cudaStreamSynchronize(p->stream));
    cleanup();
    return 0;
}

}  // extern "C"

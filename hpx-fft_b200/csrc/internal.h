// internal.h -- plan object and the launcher interface shared by the translation units of
// libhpxfft_b200.so.  The library is split so that the heavy template instantiations (row kernels,
// fused column kernel pairs) compile in parallel; every kernel is launched from the unit that defines it.
#pragma once
#include "../../include/hpxfft_b200.h"

#include "layout.cuh"

#include <nccl.h>  // types only: the library is dlopen'ed lazily (plan.cu: NcclApi)

#include <map>
#include <string>
#include <vector>

namespace hpxfft_b200 {

int fail(int code, const char *fmt, ...);
const char *last_error_string();

#define CU(call)                                                                                                     \
    do {                                                                                                             \
        cudaError_t e_ = (call);                                                                                     \
        if (e_ != cudaSuccess)                                                                                       \
            return ::hpxfft_b200::fail(HPXFFT_B200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                                       __LINE__);                                                                    \
    } while (0)

enum Mode { MODE_SHARED = 0, MODE_SCATTER = 1, MODE_ALL_TO_ALL = 2, MODE_P2P = 3 };
// how the two slab exchanges move their bytes
enum Transport {
    TR_NONE = 0,  // one rank
    TR_NCCL = 1,  // grouped ncclSend/ncclRecv + unpack kernel
    TR_CE = 2,    // copy-engine peer copies over IPC windows, chunked so that they overlap the FFT kernels
    TR_FUSED = 3  // the FFT kernels store straight into the peers' windows (no staging, no exchange phase)
};

inline bool is_pow2(size_t v) { return v && !(v & (v - 1)); }
struct BlueStage;

template <class K> int set_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024) CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
    return 0;
}

}  // namespace hpxfft_b200

struct hpxfft_b200_plan {
    int rank = 0, P = 1, device = 0, mode = hpxfft_b200::MODE_SHARED, transport = hpxfft_b200::TR_NONE;
    int sm_count = 148;
    size_t nxl = 0, n_col = 0, ny = 0, cy = 0, nx = 0, m = 0;
    // column ownership
    unsigned wq0 = 0, w = 0, c0 = 0, ntiles = 0;
    std::vector<unsigned> ntiles_of, w_of, c0_of;
    // column FFT decomposition
    unsigned n1 = 1, n2 = 1;
    unsigned col_split = 1;           // radix of the decimation-in-frequency pre-stage folded into the level-A load (nx = col_split*n1*n2)
    bool two_level = false;
    bool rows_generic = false, cols_generic = false; // lengths that are not powers of two: direct O(n^2) DFT (small sizes only)
    bool rows_mixed = false, cols_mixed = false;     // ... n = t*q, t odd: mixed radix (kernels_generic.cuh)
    unsigned gen_ct = 1, gen_cq = 0;                 // column length nx = gen_ct * gen_cq
    hpxfft_b200::cd *S1 = nullptr;                   // output of the odd-radix column pre-stage
    hpxfft_b200_plan *colsub = nullptr;              // plan-view of the power-of-two column stage (length gen_cq, reads S1)
    bool rows_blue = false, cols_blue = false;       // ... any other length above the direct-DFT bound: Bluestein (kernels_bluestein.cuh)
    hpxfft_b200::BlueStage *blue_r = nullptr, *blue_c = nullptr;
    // device buffers
    double *V = nullptr;              // slab, nxl x n_col doubles
    hpxfft_b200::cd *bufA = nullptr;  // send staging of exchange #1 and #2 (TR_NCCL, TR_CE)
    hpxfft_b200::cd *bufB = nullptr;  // I (intermediate / receive window of exchange #1); receive buffer of #2 (TR_NCCL)
    hpxfft_b200::cd *zraw = nullptr;  // un-split row spectra, only for rows longer than 32768 reals
    hpxfft_b200::cd *S = nullptr;     // four-step scratch (full array, or an L2-resident ring of strips when fused)
    bool fused = false;               // level A + level B in one persistent launch
    unsigned lag = 0, nslot = 0, fused_grid = 0;
    unsigned *ctl = nullptr;          // tile counter + per-strip completion counters
    hpxfft_b200::cd *tw_row = nullptr, *tw_col = nullptr;
    hpxfft_b200::cd *tw_il = nullptr; // inter-level twiddles of the four-step column FFT, [x2][k1] = w_nx^(k1*x2)
    size_t bytesA = 0, bytesB = 0, bytesS = 0;
    // peer windows (TR_CE, TR_FUSED)
    std::vector<void *> peerI, peerV;
    bool ipc_imported = false;
    // sub-slab chunks: rows per chunk of the row pass / strips per chunk of the column pass
    int chunks_r = 1, chunks_c = 1;
    hpxfft_b200::cd *bufC = nullptr;     // receive buffer of exchange #2 when it overlaps the column pass (TR_NCCL pipelined)
    cudaStream_t cstream = nullptr;      // high-priority communication stream (TR_NCCL pipelined)
    std::vector<cudaStream_t> pstream;   // TR_CE: one copy stream per peer
    std::vector<cudaEvent_t> ev_chunk;   // per-chunk "produced" events
    std::vector<cudaEvent_t> ev_peer;    // TR_CE: [2][P] copy-stream completion of exchange #1 / #2
    std::vector<cudaEvent_t> ev_comm;    // TR_CE: [EV_SETS][4] first/last copy of each exchange (timed)
    int sm_reserve = 0;                  // SMs left free for NCCL's kernels while the row kernel runs
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
    std::string plan_flag, row_desc, col_desc;
};

namespace hpxfft_b200 {

// ---- launchers (launch_rows.cu, launch_cols.cu, launch_fused.cu, launch_misc.cu) ---------------------
// all enqueue on p->stream; none synchronises
int launch_rows(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch, size_t m);
int rows_launch_count(size_t m);
// rows longer than one pencil (launch_rows_long.cu, launch_rows_ditc.cu) and the environment knobs they share with launch_rows.cu
int launch_rows_longer(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch, size_t m);
int launch_rows_ditc(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch, int C, bool general);
bool rows_general();
bool rows_prefetch(bool dflt);
int launch_cols(const hpxfft_b200_plan *p, const InterView &in, const ColDst &out, unsigned ntiles, cd *S, unsigned nx, unsigned n1,
                unsigned n2, bool two_level, int *launches, cudaEvent_t mid = nullptr);
int launch_cols_fused(const hpxfft_b200_plan *p, const InterView &in, const ColDst &out, unsigned ct0, unsigned ntiles);
int fused_blocks_per_sm(unsigned n1, unsigned n2, unsigned split, int *bps);
bool fused_pair_exists(unsigned n1, unsigned n2);

int launch_fill(const hpxfft_b200_plan *p, int pattern, unsigned long long seed);
int launch_unpack(const hpxfft_b200_plan *p, cudaStream_t s);
int launch_unpack_chunk(const hpxfft_b200_plan *p, const UnpackChunk &u, cudaStream_t s);
int launch_tile(const cd *A, cd *I, unsigned n, unsigned width, cudaStream_t s);
int launch_untile(const cd *I, cd *A, unsigned n, unsigned width, cudaStream_t s);
int launch_rows_generic(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch, size_t m);
int launch_cols_generic(const hpxfft_b200_plan *p, const InterView &in, const ColDst &out, unsigned ntiles, unsigned nx);

// n = t * q, t odd (launch_generic.cu)
void gen_factor(size_t n, unsigned &t, unsigned &q, unsigned &lg);
bool gen_rows_supported(size_t m);
bool gen_cols_supported(size_t nx);
int make_col_stage(const hpxfft_b200_plan *parent, unsigned q, unsigned nstrips, hpxfft_b200_plan **out);
void free_col_stage(hpxfft_b200_plan *s);
int run_col_stage(const hpxfft_b200_plan *s, const InterView &iv, const ColDst &out, unsigned nstrips, int *launches);
int gen_setup_col_stage(hpxfft_b200_plan *p);
void gen_free_col_stage(hpxfft_b200_plan *p);
int launch_rows_mixed(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch, size_t m);
int launch_cols_mixed(const hpxfft_b200_plan *p, const InterView &in, const ColDst &out, int *launches);
// Bluestein (launch_bluestein.cu): any length, on top of the power-of-two column stage
struct BlueStage;
int blue_setup(hpxfft_b200_plan *p, bool rows, size_t n, unsigned max_strips);
void blue_free(hpxfft_b200_plan *p);
int launch_rows_blue(const hpxfft_b200_plan *p, const RowDst &dst, unsigned nrows, const cd *V, unsigned pitch);
int launch_cols_blue(const hpxfft_b200_plan *p, const InterView &in, const ColDst &out, int *launches);

// exp(-2 pi i k / n), k < n, rounded from long double; exact on the axes and diagonals
void make_twiddles(std::vector<double2> &t, size_t n);

}  // namespace hpxfft_b200

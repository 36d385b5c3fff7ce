// plan.cu -- plan object and C ABI (include/hpxfft_b200.h) of libhpxfft_b200.so.
//
// Host-side counterpart of hpxfft::shared::loop::initialize / fft_2d_r2c_par
// (core/src/shared/loop.cpp:56-113,158-189) and hpxfft::distributed::loop::initialize / fft_2d_r2c
// (core/src/distributed/loop.cpp:130-347) of the reference: dimension inference, buffers, "plans"
// (twiddle tables + kernel selection), communicator, the phase sequence and its timers.
#include "internal.h"

#include <dlfcn.h>
#include <sched.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

using namespace hpxfft_b200;

namespace {

#define NC(call)                                                                                                \
    do {                                                                                                        \
        ncclResult_t r_ = (call);                                                                               \
        if (r_ != ncclSuccess)                                                                                  \
            return fail(HPXFFT_B200_ENCCL, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, \
                        __LINE__);                                                                              \
    } while (0)

// NCCL entry points, resolved at first use (a host process that already carries its own libnccl.so.2,
// e.g. PyTorch's bundled one, is reused).
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

// closes an open NCCL group on every exit path
struct NcclGroup {
    bool open = false;
    int start()
    {
        NC(g_nccl.GroupStart());
        open = true;
        return 0;
    }
    int end()
    {
        open = false;
        NC(g_nccl.GroupEnd());
        return 0;
    }
    ~NcclGroup()
    {
        if (open) g_nccl.GroupEnd();
    }
};

// [x2][k1] = w_nx^(k1*x2), nx = n1*n2, from the length-nx table
void make_interlevel(std::vector<double2> &w2, const std::vector<double2> &t, unsigned n1, unsigned n2)
{
    const size_t nx = (size_t) n1 * n2;
    w2.resize(nx);
    for (size_t x2 = 0; x2 < n2; ++x2)
        for (size_t k1 = 0; k1 < n1; ++k1) w2[x2 * n1 + k1] = t[(k1 * x2) % nx];
}

// nx = split * n1 * n2.  split = 2 (opt-in, HPXFFT_B200_COLSPLIT=1, nx = 32768 only) runs the 128 x 128 tiles of the 16384 case
// (5 CTAs per SM) behind a radix-2 pre-stage instead of 256 x 128 tiles (2 CTAs per SM).  Measured SLOWER (7.9 vs 6.5 ms per
// 32768^2: every input element is loaded twice and the second load is not served by L2), so it is not the default.
void choose_col_split(size_t nx, unsigned &n1, unsigned &n2, bool &two_level, unsigned *split = nullptr)
{
    if (split) {
        *split = 1;
        const char *e = getenv("HPXFFT_B200_COLSPLIT");
        if (nx == 32768 && e && e[0] == '1') {
            *split = 2;
            n1 = n2 = 128;
            two_level = true;
            return;
        }
    }
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

int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}

// elements of rank q's I-layout block as written by ONE source rank
unsigned long long iblock(const hpxfft_b200_plan *p, int q) { return (unsigned long long) p->ntiles_of[q] * p->nxl * CW; }

// Destination of the row pass.  Own columns always go straight into this rank's I; remote columns go into
// the peer's I (TR_FUSED) or into the send staging buffer bufA (TR_NCCL, TR_CE).
// row0: first local row of the launch (sub-slab chunks).
void fill_rowdst(const hpxfft_b200_plan *p, RowDst &d, size_t row0 = 0)
{
    d.tile_stride = (unsigned long long) p->nxl * CW;
    d.cy = (unsigned) p->cy;
    d.wq0 = p->wq0;
    d.P = (unsigned) p->P;
    unsigned long long off = 0;
    for (int q = 0; q < p->P; ++q) {
        const unsigned long long blk = iblock(p, q);
        if (q == p->rank)
            d.base[q] = p->bufB + (unsigned long long) p->rank * blk;
        else if (p->transport == TR_FUSED)
            d.base[q] = (cd *) p->peerI[q] + (unsigned long long) p->rank * blk;
        else
            d.base[q] = p->bufA + off;
        d.base[q] += (unsigned long long) row0 * CW;
        off += blk;
    }
}

void fill_coldst(const hpxfft_b200_plan *p, ColDst &d)
{
    d.nxl = (unsigned) p->nxl;
    d.shift = pow2_shift(d.nxl);
    d.w = p->w;
    d.vt = 1;
    for (int r = 0; r < p->P; ++r) {
        if (r == p->rank) {
            d.base[r] = (cd *) p->V;
            d.pitch[r] = (unsigned) p->cy;
            d.col0[r] = (int) p->c0;
        } else if (p->transport == TR_FUSED) {
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

void fill_interview(const hpxfft_b200_plan *p, InterView &iv)
{
    iv.base = p->bufB;
    iv.nxl = (unsigned) p->nxl;
    iv.shift = pow2_shift(iv.nxl);
    iv.tile_stride = (unsigned long long) p->nxl * CW;
    iv.rank_stride = (unsigned long long) p->ntiles * p->nxl * CW;
}

bool a2a_stepwise()
{
    static const int v = env_int("HPXFFT_B200_A2A_STEPWISE", 0);
    return v == 1;
}

// stream-ordered barrier across the ranks: nobody passes before everybody's earlier work on `stream` is done
int barrier_on_stream(hpxfft_b200_plan *p)
{
    NC(g_nccl.AllReduce(p->d_barrier, p->d_barrier, 1, ncclInt, ncclSum, p->comm, p->stream));
    return 0;
}

// ---- TR_NCCL ---------------------------------------------------------------------------------------------
// all_to_all: every rank's P-1 sends and receives in ONE NCCL group (rotation order), the counterpart of
//             hpx::collectives::all_to_all (core/src/distributed/loop.cpp:72-84).
// scatter   : P rooted scatters, ONE NCCL group per root issued in root order -- the literal shape of
//             scatter_to / scatter_from on P communicators (core/src/distributed/loop.cpp:39-69,158-167).  NCCL runs the
//             groups of one communicator one after the other, so only one root sends at a time.
template <class SendFn, class RecvFn> int exchange_groups(hpxfft_b200_plan *p, cudaStream_t, SendFn send, RecvFn recv)
{
    const int P = p->P, me = p->rank;
    NcclGroup g;
    if (p->mode == MODE_SCATTER) {
        for (int root = 0; root < P; ++root) {
            if (int rc = g.start()) return rc;
            if (root == me) {
                for (int s = 1; s < P; ++s)
                    if (int rc = send((me + s) % P)) return rc;
            } else if (int rc = recv(root))
                return rc;
            if (int rc = g.end()) return rc;
        }
        return 0;
    }
    const bool stepwise = a2a_stepwise();
    if (!stepwise)
        if (int rc = g.start()) return rc;
    for (int s = 1; s < P; ++s) {
        if (stepwise)
            if (int rc = g.start()) return rc;
        if (int rc = send((me + s) % P)) return rc;
        if (int rc = recv((me - s + P) % P)) return rc;
        if (stepwise)
            if (int rc = g.end()) return rc;
    }
    if (!stepwise)
        if (int rc = g.end()) return rc;
    return 0;
}

// exchange #1: block (r -> q) = I-layout tiles of rank q's columns for my rows
int exchange1_nccl(hpxfft_b200_plan *p)
{
    std::vector<unsigned long long> soff(p->P + 1, 0);
    for (int q = 0; q < p->P; ++q) soff[q + 1] = soff[q] + iblock(p, q);
    const unsigned long long rblk = iblock(p, p->rank); // what every peer sends me
    auto send = [&](int to) -> int {
        NC(g_nccl.Send(p->bufA + soff[to], (soff[to + 1] - soff[to]) * 2, ncclDouble, to, p->comm, p->stream));
        return 0;
    };
    auto recv = [&](int from) -> int {
        NC(g_nccl.Recv(p->bufB + (unsigned long long) from * rblk, rblk * 2, ncclDouble, from, p->comm, p->stream));
        return 0;
    };
    return exchange_groups(p, p->stream, send, recv);
}

// exchange #2: block (q -> r) = dense [nxl][w_q] result rows of rank r
int exchange2_nccl(hpxfft_b200_plan *p)
{
    const unsigned long long sblk = (unsigned long long) p->nxl * p->w;
    auto send = [&](int to) -> int {
        NC(g_nccl.Send(p->bufA + (unsigned long long) to * sblk, sblk * 2, ncclDouble, to, p->comm, p->stream));
        return 0;
    };
    auto recv = [&](int from) -> int {
        NC(g_nccl.Recv(p->bufB + (unsigned long long) p->nxl * p->c0_of[from], (unsigned long long) p->nxl * p->w_of[from] * 2, ncclDouble,
                       from, p->comm, p->stream));
        return 0;
    };
    return exchange_groups(p, p->stream, send, recv);
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

cudaEvent_t *event_set(hpxfft_b200_plan *p)
{
    cudaEvent_t *ev = p->evs.data() + (size_t) (p->nrec % hpxfft_b200_plan::EV_SETS) * hpxfft_b200_plan::EV_PER_SET;
    return ev;
}

// ---- TR_NCCL with sub-slab pipelining (opt-in, HPXFFT_B200_CHUNKS with HPXFFT_B200_A2A=nccl) --------------------
// the exchange of row chunk s overlaps the row FFTs of chunk s+1, the exchange (+ unpack) of strip chunk t
// overlaps the column FFTs of chunk t+1.  I is chunk-major here: [r][s][ct][js][c].
int enqueue_transform_nccl_pipelined(hpxfft_b200_plan *p)
{
    const int P = p->P, me = p->rank, Sr = p->chunks_r, Sc = p->chunks_c;
    int launches = 0;
    const size_t nxs = p->nxl / Sr;
    cudaEvent_t *ev = event_set(p);
    const long set = p->nrec % hpxfft_b200_plan::EV_SETS;
    p->nrec += 1;
    cudaEvent_t *evr = p->ev_chunk.data(), *evc = p->ev_chunk.data() + Sr;
    cudaEvent_t ev_x1 = p->ev_peer[0], ev_x2 = p->ev_peer[1];
    cudaEvent_t *evm = p->ev_comm.data() + (size_t) set * 4;

    std::vector<unsigned long long> soff(P + 1, 0);
    for (int q = 0; q < P; ++q) soff[q + 1] = soff[q] + iblock(p, q);
    const unsigned long long rblk = iblock(p, me);

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
        launches += rows_launch_count(p->m);
        CU(cudaEventRecord(evr[s], p->stream));
        CU(cudaStreamWaitEvent(p->cstream, evr[s], 0));
        if (s == 0) CU(cudaEventRecord(evm[0], p->cstream));
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
        if (int rc = exchange_groups(p, p->cstream, send, recv)) return rc;
    }
    CU(cudaEventRecord(ev[1], p->stream));
    CU(cudaEventRecord(evm[1], p->cstream));
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
            cdst.vt = 1;
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
        if (t == 0) CU(cudaEventRecord(evm[2], p->cstream));
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
        if (int rc = exchange_groups(p, p->cstream, send, recv)) return rc;
        if (any) {
            if (int rc = launch_unpack_chunk(p, u, p->cstream)) return rc;
            launches += 1;
        }
    }
    CU(cudaEventRecord(ev[3], p->stream));
    CU(cudaEventRecord(evm[3], p->cstream));
    CU(cudaEventRecord(ev_x2, p->cstream));
    CU(cudaStreamWaitEvent(p->stream, ev_x2, 0));
    CU(cudaEventRecord(ev[4], p->stream));
    CU(cudaEventRecord(ev[5], p->stream));
    p->launches = launches;
    return 0;
}

// ---- TR_CE: copy-engine all-to-all over the peers' IPC windows ---------------------------------------------------
// The FFT kernels write remote blocks into the local staging buffer bufA; cudaMemcpy2DAsync peer copies (one
// stream per peer, no SMs) push them into the owners' windows: exchange #1 into the peer's I, exchange #2
// straight into the peer's slab V (strided 2-D copy -- no unpack pass).  Rows and strips are cut into
// chunks so that the copies of chunk s run while chunk s+1 is being computed.  A 1-int NCCL all-reduce on
// the main stream is the stream-ordered barrier that tells every rank its window is complete.
int enqueue_transform_ce(hpxfft_b200_plan *p)
{
    const int P = p->P, me = p->rank, Sr = p->chunks_r, Sc = p->chunks_c;
    int launches = 0;
    const size_t nxs = p->nxl / Sr;
    cudaEvent_t *ev = event_set(p);
    const long set = p->nrec % hpxfft_b200_plan::EV_SETS;
    p->nrec += 1;
    cudaEvent_t *evr = p->ev_chunk.data(), *evc = p->ev_chunk.data() + Sr;
    cudaEvent_t *evm = p->ev_comm.data() + (size_t) set * 4;
    const int first_peer = (me + 1) % P;

    std::vector<unsigned long long> soff(P + 1, 0);
    for (int q = 0; q < P; ++q) soff[q + 1] = soff[q] + iblock(p, q);

    CU(cudaEventRecord(ev[0], p->stream));
    // ---- dimension 1
    for (int s = 0; s < Sr; ++s) {
        RowDst rd;
        fill_rowdst(p, rd, (size_t) s * nxs);
        if (int rc = launch_rows(p, rd, (unsigned) nxs, (const cd *) p->V + (size_t) s * nxs * p->cy, (unsigned) p->cy, p->m)) return rc;
        launches += rows_launch_count(p->m);
        CU(cudaEventRecord(evr[s], p->stream));
        for (int k = 1; k < P; ++k) {
            const int q = (me + k) % P;
            cudaStream_t cs = p->pstream[q];
            CU(cudaStreamWaitEvent(cs, evr[s], 0));
            if (s == 0 && q == first_peer) CU(cudaEventRecord(evm[0], cs));
            const cd *src = p->bufA + soff[q] + (unsigned long long) s * nxs * CW;
            cd *dst = (cd *) p->peerI[q] + (unsigned long long) me * iblock(p, q) + (unsigned long long) s * nxs * CW;
            if (Sr == 1)
                CU(cudaMemcpyAsync(dst, src, iblock(p, q) * sizeof(cd), cudaMemcpyDeviceToDevice, cs));
            else
                CU(cudaMemcpy2DAsync(dst, p->nxl * CW * sizeof(cd), src, p->nxl * CW * sizeof(cd), nxs * CW * sizeof(cd), p->ntiles_of[q],
                                     cudaMemcpyDeviceToDevice, cs));
        }
    }
    CU(cudaEventRecord(ev[1], p->stream));
    for (int k = 1; k < P; ++k) {
        const int q = (me + k) % P;
        CU(cudaEventRecord(p->ev_peer[q], p->pstream[q]));
        CU(cudaStreamWaitEvent(p->stream, p->ev_peer[q], 0));
    }
    CU(cudaEventRecord(evm[1], p->stream));
    if (int rc = barrier_on_stream(p)) return rc;
    CU(cudaEventRecord(ev[2], p->stream));
    CU(cudaEventRecord(ev[6], p->stream));

    // ---- dimension 2
    InterView iv;
    fill_interview(p, iv);
    ColDst cdst;
    fill_coldst(p, cdst);
    for (int t = 0; t < Sc; ++t) {
        unsigned t0, t1, col0, wc;
        chunk_bounds(p->ntiles, p->w, Sc, t, t0, t1, col0, wc);
        if (t1 > t0) {
            if (!p->fused) {
                if (int rc = launch_cols(p, iv, cdst, p->ntiles, p->S, (unsigned) p->nx, p->n1, p->n2, p->two_level, &launches, nullptr)) return rc;
            } else {
                if (int rc = launch_cols_fused(p, iv, cdst, t0, t1 - t0)) return rc;
                launches += 1;
            }
        }
        CU(cudaEventRecord(evc[t], p->stream));
        for (int k = 1; k < P; ++k) {
            const int r = (me + k) % P;
            cudaStream_t cs = p->pstream[r];
            CU(cudaStreamWaitEvent(cs, evc[t], 0));
            if (t == 0 && r == first_peer) CU(cudaEventRecord(evm[2], cs));
            if (wc == 0) continue;
            const cd *src = p->bufA + (unsigned long long) r * p->nxl * p->w + col0;
            cd *dst = (cd *) p->peerV[r] + p->c0 + col0;
            CU(cudaMemcpy2DAsync(dst, p->cy * sizeof(cd), src, (size_t) p->w * sizeof(cd), (size_t) wc * sizeof(cd), p->nxl,
                                 cudaMemcpyDeviceToDevice, cs));
        }
    }
    CU(cudaEventRecord(ev[3], p->stream));
    for (int k = 1; k < P; ++k) {
        const int q = (me + k) % P;
        CU(cudaEventRecord(p->ev_peer[P + q], p->pstream[q]));
        CU(cudaStreamWaitEvent(p->stream, p->ev_peer[P + q], 0));
    }
    CU(cudaEventRecord(evm[3], p->stream));
    if (int rc = barrier_on_stream(p)) return rc;
    CU(cudaEventRecord(ev[4], p->stream));
    CU(cudaEventRecord(ev[5], p->stream));
    p->launches = launches;
    return 0;
}

int enqueue_transform(hpxfft_b200_plan *p)
{
    if (p->P > 1 && (p->transport == TR_FUSED || p->transport == TR_CE) && !p->ipc_imported)
        return fail(HPXFFT_B200_ESTATE, "plan uses peer windows: hpxfft_b200_ipc_export / hpxfft_b200_ipc_import have not been called");
    if (p->transport == TR_CE) return enqueue_transform_ce(p);
    if (p->transport == TR_NCCL && (p->chunks_r > 1 || p->chunks_c > 1)) return enqueue_transform_nccl_pipelined(p);
    int launches = 0;
    RowDst rd;
    fill_rowdst(p, rd);
    ColDst cdst;
    fill_coldst(p, cdst);
    InterView iv;
    fill_interview(p, iv);

    cudaEvent_t *ev = event_set(p);
    p->nrec += 1;
    CU(cudaEventRecord(ev[0], p->stream));
    // phase 1: r2c rows (+ fused split / transpose)           -> first_fftw (first_split, first_trans fused)
    if (int rc = launch_rows(p, rd, (unsigned) p->nxl, (const cd *) p->V, (unsigned) p->cy, p->m)) return rc;
    launches += rows_launch_count(p->m);
    CU(cudaEventRecord(ev[1], p->stream));
    // phase 2: exchange #1                                    -> first_comm
    if (p->transport == TR_FUSED) {
        if (int rc = barrier_on_stream(p)) return rc;
    } else if (p->transport == TR_NCCL) {
        if (int rc = exchange1_nccl(p)) return rc;
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
    if (p->transport == TR_FUSED) {
        if (int rc = barrier_on_stream(p)) return rc;
    } else if (p->transport == TR_NCCL) {
        if (int rc = exchange2_nccl(p)) return rc;
    }
    CU(cudaEventRecord(ev[4], p->stream));
    // phase 5: unpack into the slab                           -> second_trans
    if (p->transport == TR_NCCL) {
        if (int rc = launch_unpack(p, p->stream)) return rc;
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
    const bool chunked = p->transport == TR_CE || (p->transport == TR_NCCL && (p->chunks_r > 1 || p->chunks_c > 1));
    double acc[10] = {0};
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
        if (chunked) {
            cudaEvent_t *evm = p->ev_comm.data() + (size_t) slot * 4;
            CU(cudaEventElapsedTime(&ms, evm[0], evm[1]));
            acc[8] += ms; // first copy of exchange #1 issued -> last one complete
            CU(cudaEventElapsedTime(&ms, evm[2], evm[3]));
            acc[9] += ms;
        }
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
    // span of the exchange traffic itself (chunked transports overlap it with the kernels); for the serial
    // transports the span IS the exposed phase
    m["first_comm_span"] = chunked ? acc[8] * sc : acc[1] * sc;
    m["second_comm_span"] = chunked ? acc[9] * sc : acc[3] * sc;
    m["timer_samples"] = (double) nset;
    return 0;
}

// CPUs of the NUMA node the device hangs off (/sys/bus/pci/devices/<id>/local_cpulist)
bool device_cpulist(int device, cpu_set_t *set)
{
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    for (char *c = bus; *c; ++c) *c = (char) tolower(*c);
    std::ifstream f(std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist");
    std::string s;
    if (!f || !std::getline(f, s) || s.empty()) return false;
    CPU_ZERO(set);
    int n = 0;
    const char *c = s.c_str();
    while (*c) {
        char *e = nullptr;
        long a = strtol(c, &e, 10), b = a;
        if (e == c) break;
        c = e;
        if (*c == '-') {
            b = strtol(c + 1, &e, 10);
            c = e;
        }
        for (long i = a; i <= b && i < CPU_SETSIZE; ++i) {
            CPU_SET((int) i, set);
            ++n;
        }
        if (*c == ',') ++c;
    }
    return n > 0;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int hpxfft_b200_version(void) { return HPXFFT_B200_VERSION; }
const char *hpxfft_b200_last_error(void) { return last_error_string(); }

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

int hpxfft_b200_bind_host_to_device(int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(HPXFFT_B200_ECUDA, "no CUDA device available (libhpxfft_b200 has no CPU fallback)");
    }
    if (device < 0) CU(cudaGetDevice(&device));
    cpu_set_t set;
    if (!device_cpulist(device, &set)) return fail(HPXFFT_B200_EINVAL, "no local_cpulist for device %d", device);
    if (sched_setaffinity(0, sizeof(set), &set) != 0) return fail(HPXFFT_B200_EINVAL, "sched_setaffinity failed for device %d", device);
    return 0;
}

void hpxfft_b200_destroy(hpxfft_b200_plan *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    for (auto &s : p->pstream)
        if (s) {
            cudaStreamSynchronize(s);
            cudaStreamDestroy(s);
        }
    if (p->ipc_imported) {
        for (int q = 0; q < p->P; ++q) {
            if (q == p->rank) continue;
            if (p->peerI[q]) cudaIpcCloseMemHandle(p->peerI[q]);
            if (p->peerV[q]) cudaIpcCloseMemHandle(p->peerV[q]);
        }
    }
    if (p->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(p->comm);
    for (auto *vec : {&p->ev_chunk, &p->ev_peer, &p->ev_comm, &p->evs})
        for (auto &e : *vec)
            if (e) cudaEventDestroy(e);
    if (p->cstream) {
        cudaStreamSynchronize(p->cstream);
        cudaStreamDestroy(p->cstream);
    }
    gen_free_col_stage(p);
    blue_free(p);
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
    cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
    // transport of the slab exchanges
    if (nranks == 1)
        p->transport = TR_NONE;
    else if (mode == MODE_P2P)
        p->transport = TR_FUSED;
    else if (mode == MODE_SCATTER)
        p->transport = TR_NCCL;
    else {
        // all_to_all: the fastest correct transport for the slab at hand (measured on 8 x B200, profiles/r2_n8_bench_*.json):
        // peer stores fused into the FFT kernels (3.56 ms per 32768^2 vs 4.90 copy-engine, 6.53 NCCL) -- unless a slab row is so long
        // that consecutive destination rows of the column kernel's 256-byte peer stores fall into different 2 MB pages
        // (131072^2: 1188 ms fused vs 98.8 ms copy-engine), where the chunked copy-engine exchange is the default.
        const char *e = getenv("HPXFFT_B200_A2A");
        if (e && !strcmp(e, "nccl")) p->transport = TR_NCCL;
        else if (e && !strcmp(e, "fused")) p->transport = TR_FUSED;
        else if (e && !strcmp(e, "ce")) p->transport = TR_CE;
        else p->transport = (n_col / 2) * sizeof(cd) <= (size_t) 512 * 1024 ? TR_FUSED : TR_CE;
    }
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
    // powers of two take the Stockham kernels; n = t * q with a small odd factor t takes the mixed-radix kernels
    // (kernels_generic.cuh; FFTW accepts every length, the reference's default example is 8 x 14 and its weak-scaling sweep
    // runs 512 * t, t = 1..32); anything else falls back to the direct-DFT kernels, bounded to sizes where O(n^2) is sane
    constexpr size_t GENERIC_MAX = 8192;
    if (p->ny < 2) return bail(fail(HPXFFT_B200_EINVAL, "unsupported ny=%zu", p->ny));
    constexpr size_t BLUE_MAX = 131072; // convolution length 2^18
    if (!is_pow2(p->m)) {
        p->rows_mixed = gen_rows_supported(p->m);
        p->rows_generic = !p->rows_mixed && p->ny <= GENERIC_MAX;
        p->rows_blue = !p->rows_mixed && !p->rows_generic; // large prime factor: Bluestein over the half-length complex row
    }
    if (!is_pow2(p->nx)) {
        p->cols_mixed = gen_cols_supported(p->nx);
        p->cols_generic = !p->cols_mixed && p->nx <= GENERIC_MAX;
        p->cols_blue = !p->cols_mixed && !p->cols_generic;
    }
    if ((is_pow2(p->m) && p->m > 65536) || (p->rows_blue && p->m > BLUE_MAX))
        return bail(fail(HPXFFT_B200_EINVAL, "unsupported ny=%zu: ny/2 must be a power of two <= 65536 or any other length <= %zu", p->ny, BLUE_MAX));
    if ((is_pow2(p->nx) && p->nx > (1u << 18)) || (p->cols_blue && p->nx > BLUE_MAX))
        return bail(fail(HPXFFT_B200_EINVAL, "unsupported nx=%zu: must be a power of two <= 2^18 or any other length <= %zu", p->nx, BLUE_MAX));
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
    choose_col_split(p->nx, p->n1, p->n2, p->two_level, &p->col_split);
    if (p->cols_generic || p->cols_mixed || p->cols_blue) {
        p->two_level = false;
        p->n1 = (unsigned) p->nx;
        p->n2 = 1;
        p->col_split = 1;
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
    p->bytesB = (size_t) p->ntiles * p->nx * CW * sizeof(cd); // I: [r][ct][j][c], also >= nxl*cy for exchange #2
    if (p->bytesB < p->nxl * p->cy * sizeof(cd)) p->bytesB = p->nxl * p->cy * sizeof(cd);
    p->bytesS = p->two_level ? (size_t) p->ntiles * p->nx * CW * sizeof(cd) : 0;
    if (p->two_level) {
        const char *e = getenv("HPXFFT_B200_FUSED");
        p->fused = !(e && e[0] == '0') && fused_pair_exists(p->n1, p->n2);
        if (!p->fused && p->col_split > 1) { // the pre-stage only exists in the fused kernel
            p->col_split = 1;
            choose_col_split(p->nx, p->n1, p->n2, p->two_level);
        }
    }
    const bool nccl_pipelined = p->transport == TR_NCCL && mode == MODE_ALL_TO_ALL && env_int("HPXFFT_B200_CHUNKS", 1) > 1;
    if (nccl_pipelined) {
        // Sub-slab pipelining of the NCCL exchanges is OFF unless asked for: NCCL's copy kernels need >= 32 CTAs
        // for full NVLink rate, which the persistent FFT kernels cannot spare.
        p->sm_reserve = 16;
        const int v = env_int("HPXFFT_B200_SM_RESERVE", 16);
        if (v >= 0 && v <= 64) p->sm_reserve = v;
    }
#define CUB(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return bail(fail(HPXFFT_B200_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)));     \
    } while (0)
    if (p->fused) {
        int bps = 1;
        if (int rc = fused_blocks_per_sm(p->n1, p->n2, p->col_split, &bps)) return bail(rc);
        {
            const int v = env_int("HPXFFT_B200_FUSED_BPS", 0);
            if (v >= 1 && v < bps) bps = v;
        }
        p->fused_grid = (unsigned) (bps * (p->sm_count - p->sm_reserve));
        const unsigned per_group = p->n1 + p->n2;
        p->lag = (unsigned) ((3 * (size_t) p->fused_grid / 2 + per_group - 1) / per_group) + 1;
        {
            const int v = env_int("HPXFFT_B200_LAG", 0);
            if (v >= 1) p->lag = (unsigned) v;
        }
        p->nslot = 2 * p->lag + 1;
        {
            const int v = env_int("HPXFFT_B200_NSLOT", 0);
            if (v > (int) p->lag) p->nslot = (unsigned) v;
        }
        if (p->nslot > p->ntiles * p->col_split) p->nslot = p->ntiles > 0 ? p->ntiles * p->col_split : 1;
        p->bytesS = (size_t) p->nslot * (p->nx / p->col_split) * CW * sizeof(cd);
    }
    if (p->transport == TR_NCCL || p->transport == TR_CE) {
        size_t tiles_all = 0;
        for (int q = 0; q < nranks; ++q) tiles_all += p->ntiles_of[q];
        p->bytesA = tiles_all * p->nxl * CW * sizeof(cd);
        const size_t ex2 = (size_t) nranks * p->nxl * p->w * sizeof(cd);
        if (p->bytesA < ex2) p->bytesA = ex2;
    }
    CUB(cudaMalloc(&p->V, bytesV));
    CUB(cudaMalloc(&p->bufB, p->bytesB));
    if (is_pow2(p->m) && p->m > 8192) // long rows: per-CTA scratch of the row kernel (one CTA per SM, rewritten per row, L2-resident)
        CUB(cudaMalloc(&p->zraw, (size_t) p->sm_count * p->m * sizeof(cd)));
    if (p->bytesS) CUB(cudaMalloc(&p->S, p->bytesS));
    if (p->fused) CUB(cudaMalloc(&p->ctl, (1 + 2 * (size_t) p->ntiles * p->col_split) * sizeof(unsigned)));
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
            if (p->col_split > 1) make_twiddles(t, p->nx / p->col_split); // inter-level factors of the length-n' transforms
            make_interlevel(w2, t, p->n1, p->n2);
            CUB(cudaMalloc(&p->tw_il, w2.size() * sizeof(double2)));
            CUB(cudaMemcpyAsync(p->tw_il, w2.data(), w2.size() * sizeof(double2), cudaMemcpyHostToDevice, p->stream));
            CUB(cudaStreamSynchronize(p->stream));
        }
    }

    if (p->cols_mixed)
        if (int rc = gen_setup_col_stage(p)) return bail(rc);
    if (p->rows_blue)
        if (int rc = blue_setup(p, true, p->m, (unsigned) ((p->nxl + CW - 1) / CW))) return bail(rc);
    if (p->cols_blue)
        if (int rc = blue_setup(p, false, p->nx, p->ntiles)) return bail(rc);

    // sub-slab chunks (rank-invariant: derived from quantities every rank computes identically)
    if (p->transport == TR_CE || nccl_pipelined) {
        const int want = env_int("HPXFFT_B200_CHUNKS", p->transport == TR_CE ? 4 : 1);
        int sr = want < 1 ? 1 : want, sc = sr;
        while (sr > 1 && (p->nxl % sr != 0 || p->nxl / sr < 8)) sr /= 2;
        if (!p->fused) sc = 1;
        unsigned min_tiles = p->ntiles_of[0]; // every rank but the last owns ntiles_of[0] strips; the last one at least as many
        while (sc > 1 && min_tiles / sc < 2 * (p->lag + 1)) sc /= 2;
        p->chunks_r = sr < 1 ? 1 : sr;
        p->chunks_c = sc < 1 ? 1 : sc;
        p->ev_chunk.assign((size_t) p->chunks_r + p->chunks_c, nullptr);
        for (auto &e : p->ev_chunk) CUB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        p->ev_comm.assign((size_t) hpxfft_b200_plan::EV_SETS * 4, nullptr);
        for (auto &e : p->ev_comm) CUB(cudaEventCreate(&e));
        p->ev_peer.assign(2 * (size_t) nranks, nullptr);
        for (auto &e : p->ev_peer) CUB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (p->transport == TR_CE) {
            p->pstream.assign(nranks, nullptr);
            for (int q = 0; q < nranks; ++q)
                if (q != rank) CUB(cudaStreamCreateWithPriority(&p->pstream[q], cudaStreamNonBlocking, hi));
        } else {
            CUB(cudaMalloc(&p->bufC, p->nxl * p->cy * sizeof(cd)));
            CUB(cudaStreamCreateWithPriority(&p->cstream, cudaStreamNonBlocking, hi));
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

    char buf[640];
    snprintf(buf, sizeof(buf), "r2c rows: n=%zu via half-length complex Stockham m=%zu (%s), %d points/thread, paired radix-16 last pass + Hermitian split",
             p->ny, p->m, p->m <= 16 ? "register-resident" : "shared-memory pencil", p->m <= 16 ? (int) p->m : 32);
    p->row_desc = buf;
    if (p->m == 8192)
        p->row_desc = "r2c rows: n=16384 via half-length complex m=8192 = 16 x 512: swizzled shared-memory pencil, warp-local in-place DIF passes "
                      "(radix 16, radix 32), paired radix-16 last pass + Hermitian split, cp.async refill of the pencil (rows_r2c_v2_kernel)";
    else if (p->m > 8192) {
        snprintf(buf, sizeof(buf), "r2c rows: n=%zu via half-length complex m=%zu = %zu x 8192: one persistent CTA per row runs %zu 8192-point "
                 "sub-transforms through one shared-memory pencil, sub-spectra parked per CTA in L2, Hermitian split in the combine "
                 "(decimation in time over sample classes: rows_dit2_kernel / rows_ditc_kernel; decimation in frequency: rows_long2_kernel / "
                 "rows_long_kernel -- chosen per launch, launch_rows_long.cu)", p->ny, p->m, p->m / 8192, p->m / 8192);
        p->row_desc = buf;
    }
    if (p->rows_generic) p->row_desc = "r2c rows: direct DFT (length is not a power of two)";
    if (p->rows_mixed) {
        unsigned t, q, lg;
        gen_factor(p->m, t, q, lg);
        snprintf(buf, sizeof(buf), "r2c rows: n=%zu via half-length complex mixed radix m = %u x %u (in-place radix-4 DIF of the stride-%u sub-sequences, radix-%u combine, Hermitian split)",
                 p->ny, t, q, t, t);
        p->row_desc = buf;
    }
    if (p->rows_blue) p->row_desc = "r2c rows: Bluestein chirp-z over the half-length complex row (two power-of-two transforms on 16-row tiles), Hermitian split";
    if (p->cols_blue)
        snprintf(buf, sizeof(buf), "c2c columns: n=%zu Bluestein chirp-z (two power-of-two column transforms of the padded convolution length)", p->nx);
    else if (p->cols_mixed)
        snprintf(buf, sizeof(buf), "c2c columns: n=%zu mixed radix %u x %u (odd-radix direct DFT pre-stage, then the power-of-two column kernels on %u virtual strips per strip)",
                 p->nx, p->gen_ct, p->gen_cq, p->gen_ct);
    else if (p->cols_generic)
        snprintf(buf, sizeof(buf), "c2c columns: n=%zu direct DFT (length is not a power of two)", p->nx);
    else if (p->two_level)
        snprintf(buf, sizeof(buf), "c2c columns: n=%zu %sfour-step %u x %u on %d-column tiles (level A strided + twiddle, level B contiguous)%s",
                 p->nx, p->col_split > 1 ? "radix-2 DIF pre-stage + " : "", p->n1, p->n2, CW,
                 p->fused ? ", fused persistent launch with L2-resident scratch ring" : "");
    else
        snprintf(buf, sizeof(buf), "c2c columns: n=%zu single Stockham tile FFT on %d-column tiles", p->nx, CW);
    p->col_desc = buf;
    *out = p;
    return 0;
}

int hpxfft_b200_ipc_count(const hpxfft_b200_plan *p)
{
    return (p && p->P > 1 && (p->transport == TR_FUSED || p->transport == TR_CE)) ? 2 : 0;
}

int hpxfft_b200_ipc_export(hpxfft_b200_plan *p, void *handles_out)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == HPXFFT_B200_IPC_HANDLE_BYTES, "ipc handle size");
    if (!p || !handles_out) return fail(HPXFFT_B200_EINVAL, "NULL argument");
    if (hpxfft_b200_ipc_count(p) == 0) return fail(HPXFFT_B200_ESTATE, "this plan exports no peer windows");
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
    if (hpxfft_b200_ipc_count(p) == 0) return fail(HPXFFT_B200_ESTATE, "this plan uses no peer windows");
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

const char *hpxfft_b200_transport(const hpxfft_b200_plan *p)
{
    if (!p) return "";
    switch (p->transport) {
    case TR_NCCL: return p->mode == MODE_SCATTER ? "nccl-rooted" : ((p->chunks_r > 1 || p->chunks_c > 1) ? "nccl-pipelined" : "nccl");
    case TR_CE: return "copy-engine";
    case TR_FUSED: return "fused-peer-store";
    default: return "none";
    }
}

int hpxfft_b200_upload(hpxfft_b200_plan *p, const double *host_slab)
{
    if (!p || !host_slab) return fail(HPXFFT_B200_EINVAL, "NULL argument");
    CU(cudaSetDevice(p->device));
    cudaEvent_t a = p->ev_io[0], b = p->ev_io[1];
    CU(cudaEventRecord(a, p->stream));
    CU(cudaMemcpyAsync(p->V, host_slab, p->nxl * p->n_col * sizeof(double), cudaMemcpyDefault, p->stream));
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
    CU(cudaMemcpyAsync(host_slab, p->V, p->nxl * p->n_col * sizeof(double), cudaMemcpyDefault, p->stream));
    CU(cudaEventRecord(b, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, a, b));
    p->meas["d2h"] = ms * 1e-3;
    return 0;
}

int hpxfft_b200_download_tile(hpxfft_b200_plan *p, size_t row0, size_t nrows, size_t col0, size_t ncols, double *host_out)
{
    if (!p || !host_out) return fail(HPXFFT_B200_EINVAL, "NULL argument");
    if (row0 + nrows > p->nxl || col0 + ncols > p->n_col || nrows == 0 || ncols == 0)
        return fail(HPXFFT_B200_EINVAL, "tile [%zu+%zu) x [%zu+%zu) outside the %zu x %zu slab", row0, nrows, col0, ncols, p->nxl, p->n_col);
    CU(cudaSetDevice(p->device));
    CU(cudaMemcpy2DAsync(host_out, ncols * sizeof(double), p->V + row0 * p->n_col + col0, p->n_col * sizeof(double), ncols * sizeof(double),
                         nrows, cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    return 0;
}

int hpxfft_b200_fill(hpxfft_b200_plan *p, int pattern, uint64_t seed)
{
    if (!p) return fail(HPXFFT_B200_EINVAL, "NULL plan");
    if (pattern < 0 || pattern > 2) return fail(HPXFFT_B200_EINVAL, "unknown pattern %d", pattern);
    CU(cudaSetDevice(p->device));
    if (int rc = launch_fill(p, pattern, seed)) return rc;
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

int hpxfft_b200_bench_exchange(hpxfft_b200_plan *p, int which, int reps, double *ms_out)
{
    if (!p || !ms_out || reps < 1 || (which != 1 && which != 2)) return fail(HPXFFT_B200_EINVAL, "bad arguments");
    if (p->transport != TR_NCCL && p->transport != TR_CE) return fail(HPXFFT_B200_ESTATE, "transport %s has no separable exchange", hpxfft_b200_transport(p));
    if (p->transport == TR_CE && !p->ipc_imported) return fail(HPXFFT_B200_ESTATE, "peer windows not imported");
    CU(cudaSetDevice(p->device));
    const int P = p->P, me = p->rank;
    std::vector<unsigned long long> soff(P + 1, 0);
    for (int q = 0; q < P; ++q) soff[q + 1] = soff[q] + iblock(p, q);
    cudaEvent_t a = p->ev_io[0], b = p->ev_io[1];
    auto once = [&]() -> int {
        if (p->transport == TR_NCCL) return which == 1 ? exchange1_nccl(p) : exchange2_nccl(p);
        CU(cudaEventRecord(p->ev_chunk[0], p->stream));
        for (int k = 1; k < P; ++k) {
            const int q = (me + k) % P;
            cudaStream_t cs = p->pstream[q];
            CU(cudaStreamWaitEvent(cs, p->ev_chunk[0], 0));
            if (which == 1)
                CU(cudaMemcpyAsync((cd *) p->peerI[q] + (unsigned long long) me * iblock(p, q), p->bufA + soff[q], iblock(p, q) * sizeof(cd),
                                   cudaMemcpyDeviceToDevice, cs));
            else
                CU(cudaMemcpy2DAsync((cd *) p->peerV[q] + p->c0, p->cy * sizeof(cd), p->bufA + (unsigned long long) q * p->nxl * p->w,
                                     (size_t) p->w * sizeof(cd), (size_t) p->w * sizeof(cd), p->nxl, cudaMemcpyDeviceToDevice, cs));
            CU(cudaEventRecord(p->ev_peer[q], cs));
            CU(cudaStreamWaitEvent(p->stream, p->ev_peer[q], 0));
        }
        return barrier_on_stream(p);
    };
    if (int rc = once()) return rc; // warm-up
    if (p->transport == TR_CE)
        if (int rc = barrier_on_stream(p)) return rc;
    CU(cudaEventRecord(a, p->stream));
    for (int i = 0; i < reps; ++i)
        if (int rc = once()) return rc;
    CU(cudaEventRecord(b, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, a, b));
    *ms_out = (double) ms / reps;
    return 0;
}

int hpxfft_b200_upload_async(hpxfft_b200_plan *p, const double *host_slab)
{
    if (!p || !host_slab) return fail(HPXFFT_B200_EINVAL, "NULL argument");
    CU(cudaSetDevice(p->device));
    CU(cudaMemcpyAsync(p->V, host_slab, p->nxl * p->n_col * sizeof(double), cudaMemcpyDefault, p->stream));
    return 0;
}

int hpxfft_b200_download_async(hpxfft_b200_plan *p, double *host_slab)
{
    if (!p || !host_slab) return fail(HPXFFT_B200_EINVAL, "NULL argument");
    CU(cudaSetDevice(p->device));
    CU(cudaMemcpyAsync(host_slab, p->V, p->nxl * p->n_col * sizeof(double), cudaMemcpyDefault, p->stream));
    return 0;
}

int hpxfft_b200_on_complete(hpxfft_b200_plan *p, hpxfft_b200_callback fn, void *user)
{
    if (!p || !fn) return fail(HPXFFT_B200_EINVAL, "NULL argument");
    CU(cudaSetDevice(p->device));
    CU(cudaLaunchHostFunc(p->stream, fn, user));
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
    fprintf(f, "FFTW c2c 1D plan:\n(hpxfft_b200 sm_100a %s)\n\n", p->col_desc.c_str());
    fclose(f);
    return 0;
}

void *hpxfft_b200_device_ptr(hpxfft_b200_plan *p) { return p ? p->V : nullptr; }
void *hpxfft_b200_stream(hpxfft_b200_plan *p) { return p ? (void *) p->stream : nullptr; }
int hpxfft_b200_launches_per_execute(const hpxfft_b200_plan *p)
{
    if (!p) return 0;
    if (p->launches > 0) return p->launches; // counted by the last execute
    int n = rows_launch_count(p->m) + ((p->two_level && !p->fused) ? 2 : 1);
    if (p->transport == TR_NCCL) n += 1;
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
    cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
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
    if (m > 8192) CUR(cudaMalloc(&p->zraw, (size_t) p->sm_count * m * sizeof(cd)));
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
    rc = launch_untile(p->bufB, (cd *) p->V, (unsigned) batch, (unsigned) cy, p->stream);
    if (rc) {
        cleanup();
        return rc;
    }
    CUR(cudaMemcpyAsync(host_rows, p->V, bytes, cudaMemcpyDeviceToHost, p->stream));
    CUR(cudaStreamSynchronize(p->stream));
    cleanup();
    return 0;
}

// variant: 0 = the kernels a plan of this length runs (fused persistent four-step when n > 256),
//          1 = the unfused kernels (single tile / level A + level B launches)
int hpxfft_b200_c2c_cols_variant(double *host_data, size_t n, size_t width, int device, int variant)
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
    cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
    p->nxl = p->nx = n;
    p->cy = width;
    p->w = (unsigned) width;
    p->ntiles = (unsigned) ((width + CW - 1) / CW);
    choose_col_split(n, p->n1, p->n2, p->two_level, variant == 0 ? &p->col_split : nullptr);
    p->fused = variant == 0 && p->two_level && fused_pair_exists(p->n1, p->n2);
    const size_t bytes = n * width * sizeof(cd), tbytes = (size_t) p->ntiles * n * CW * sizeof(cd);
    cd *A = nullptr;
    std::vector<double2> t;
    make_twiddles(t, n);
    int rc = 0;
    auto cleanup = [&]() {
        cudaFree(A);
        cudaFree(p->bufB);
        cudaFree(p->S);
        cudaFree(p->ctl);
        cudaFree(p->tw_col);
        cudaFree(p->tw_il);
        if (p->stream) cudaStreamDestroy(p->stream);
        p->bufB = nullptr; p->S = nullptr; p->ctl = nullptr; p->tw_col = nullptr; p->tw_il = nullptr; p->stream = nullptr;
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
    size_t sbytes = tbytes;
    if (p->fused) {
        int bps = 1;
        if ((rc = fused_blocks_per_sm(p->n1, p->n2, p->col_split, &bps))) {
            cleanup();
            return rc;
        }
        p->fused_grid = (unsigned) (bps * p->sm_count);
        const unsigned per_group = p->n1 + p->n2;
        p->lag = (unsigned) ((3 * (size_t) p->fused_grid / 2 + per_group - 1) / per_group) + 1;
        p->nslot = 2 * p->lag + 1;
        if (p->nslot > p->ntiles * p->col_split) p->nslot = p->ntiles * p->col_split;
        sbytes = (size_t) p->nslot * (n / p->col_split) * CW * sizeof(cd);
        CUC(cudaMalloc(&p->ctl, (1 + 2 * (size_t) p->ntiles * p->col_split) * sizeof(unsigned)));
    }
    if (p->two_level) CUC(cudaMalloc(&p->S, sbytes));
    CUC(cudaMalloc(&p->tw_col, t.size() * sizeof(double2)));
    CUC(cudaMemcpyAsync(p->tw_col, t.data(), t.size() * sizeof(double2), cudaMemcpyHostToDevice, p->stream));
    std::vector<double2> w2;
    if (p->two_level) {
        if (p->col_split > 1) make_twiddles(t, n / p->col_split);
        make_interlevel(w2, t, p->n1, p->n2);
        CUC(cudaMalloc(&p->tw_il, w2.size() * sizeof(double2)));
        CUC(cudaMemcpyAsync(p->tw_il, w2.data(), w2.size() * sizeof(double2), cudaMemcpyHostToDevice, p->stream));
    }
    CUC(cudaMemcpyAsync(A, host_data, bytes, cudaMemcpyHostToDevice, p->stream));
    if ((rc = launch_tile(A, p->bufB, (unsigned) n, (unsigned) width, p->stream))) {
        cleanup();
        return rc;
    }
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
    cdst.vt = 1;
    cdst.base[0] = A;
    cdst.pitch[0] = (unsigned) width;
    cdst.col0[0] = 0;
    if (p->fused)
        rc = launch_cols_fused(p, iv, cdst, 0u, p->ntiles);
    else
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

int hpxfft_b200_c2c_cols(double *host_data, size_t n, size_t width, int device)
{
    return hpxfft_b200_c2c_cols_variant(host_data, n, width, device, 0);
}

}  // extern "C"

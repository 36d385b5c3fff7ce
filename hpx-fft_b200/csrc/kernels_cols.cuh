// kernels_cols.cuh -- forward c2c FFT along x (the strided axis) on column tiles.
//
// Replaces fft_1d_c2c_inplace on the transposed array plus both local transposes of the reference
// (core/src/shared/loop.cpp:11-15,18-25,46-53; core/src/distributed/loop.cpp:12-16,87-127):
// instead of transposing so that x becomes contiguous, a CTA owns a tile of CW = 16 adjacent ky
// columns and runs the FFT down the rows.  All global accesses are 256-byte segments, the shared
// memory tile is [point][column] with the column index fastest across threads, which makes every
// shared-memory access conflict-free without padding.  Twiddles never come from global memory inside
// the butterfly loops: the per-pass factors live in a small shared-memory table built once per CTA and
// the inter-level factors w_nx^(k1*x2) are one contiguous row of a precomputed [x2][k1] table.
//
// nx <= 256        : one Stockham FFT per tile (cols_single_kernel).
// nx = n1*n2 > 256 : four-step.  Level A: for every x2, FFT over x1 (stride n2 rows), multiply by
//                    w_nx^(k1*x2), write scratch S[strip][k1][x2][c].  Level B: for every k1, FFT over x2
//                    (contiguous in S), result row kx = k1 + n1*k2 goes straight to its final position.
//                    Run as two launches (cols_levelA/B_kernel, S = full array in HBM) or fused in one
//                    persistent launch whose scratch ring stays in L2 (cols_fused_kernel).
#pragma once
#include "layout.cuh"

namespace hpxfft_b200 {

__host__ __device__ constexpr int col_npass(int N) { return N <= 16 ? 1 : (N <= 256 ? 2 : 3); }
__host__ __device__ constexpr int col_radix(int N, int p)
{
    if (N <= 16) return N;
    if (N == 32) return p == 0 ? 8 : 4;
    if (N == 64) return 8;
    if (N == 128) return p == 0 ? 16 : 8;
    if (N == 256) return 16;
    return 8; // 512 = 8*8*8
}
__host__ __device__ constexpr int col_pt(int N) { return N < 16 ? N : 16; }       // points per thread
__host__ __device__ constexpr int col_threads(int N) { return (N / col_pt(N)) * CW; }
// shared-memory twiddle table: pass p >= 1 holds w_{NS*R}^(r*k) at [k*R + r], k < NS (= product of earlier radices)
__host__ __device__ constexpr int col_tw_entries(int N)
{
    return col_npass(N) == 1 ? 0 : (col_npass(N) == 2 ? N : col_radix(N, 0) * col_radix(N, 1) + N);
}
__host__ __device__ constexpr size_t col_tile_bytes(int N) { return N <= 16 ? 0 : (size_t) N * CW * sizeof(cd); }
__host__ __device__ constexpr size_t col_tw_bytes(int N) { return (size_t) col_tw_entries(N) * sizeof(cd); }

// CTA-wide barrier (BAR_THREADS == 0) or named barrier 1 over the first BAR_THREADS threads (warp-
// specialised kernels whose producer warp must not take part)
template <int BAR_THREADS> __device__ __forceinline__ void tile_barrier()
{
    if constexpr (BAR_THREADS == 0)
        __syncthreads();
    else
        asm volatile("bar.sync 1, %0;" ::"n"(BAR_THREADS) : "memory");
}

struct NoHook {
    __device__ __forceinline__ void operator()() const {}
};

// Fills the per-pass twiddle table of a length-N tile FFT from the global table w_L^i (L = N * tws).
// Call with all `nthreads` participating threads, then barrier before the first tile_fft.
template <int N> __device__ __forceinline__ void fill_pass_twiddles(cd *ptw, const cd *__restrict__ tw, unsigned tws, int tid, int nthreads)
{
    constexpr int NP = col_npass(N);
    if constexpr (NP >= 2) {
        constexpr int R0 = col_radix(N, 0), R1 = col_radix(N, 1);
        for (int i = tid; i < R0 * R1; i += nthreads) {
            const int k = i / R1, r = i - k * R1;
            ptw[i] = ldtw(tw, (unsigned) (r * k * (N / (R0 * R1))) * tws);
        }
        if constexpr (NP == 3) {
            constexpr int R2 = col_radix(N, 2);
            for (int i = tid; i < N; i += nthreads) {
                const int k = i / R2, r = i - k * R2;
                ptw[R0 * R1 + i] = ldtw(tw, (unsigned) (r * k) * tws); // NS*R == N: stride 1 in the length-N table
            }
        }
    }
}

template <int N, int PT, int R, int NS, bool FIRST, bool LAST, int BAR_THREADS, class LD, class ST, class HOOK, class HOOK2>
__device__ __forceinline__ void col_pass(cd (&v)[PT], cd *smem, const cd *ptw, LD &ld, ST &st, int u, int c, bool active,
                                         HOOK &after_load, HOOK2 &tile_dead)
{
    constexpr int NB = PT / R, T = N / R, U = N / PT;
    constexpr int LGR = ilog2(R);
    if (active) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int j = u + b * U;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (FIRST)
                    v[b * R + r] = ld(j + r * T, c);
                else
                    v[b * R + r] = smem[(j + r * T) * CW + c];
            }
        }
    }
    if (FIRST) after_load();                          // first-pass inputs are in registers
    if (LAST) tile_dead();                            // last read of the tile buffer (non-blocking hook)
    if (!FIRST && !LAST) tile_barrier<BAR_THREADS>(); // every thread has read its inputs before the in-place overwrite
    if (active) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int j = u + b * U;
            const int k = j & (NS - 1);
            cd w[R];
#pragma unroll
            for (int r = 0; r < R; ++r) w[r] = v[b * R + r];
            if (NS > 1) {
#pragma unroll
                for (int r = 1; r < R; ++r) w[r] = cmul(w[r], ptw[k * R + r]);
            }
            fft_dif<R>(w);
            const int j0 = ((j - k) << LGR) + k; // (j / NS) * NS * R + k
#pragma unroll
            for (int s = 0; s < R; ++s) {
                const int o = j0 + s * NS;
                if (LAST)
                    st(o, c, w[bitrev(s, LGR)]);
                else
                    smem[o * CW + c] = w[bitrev(s, LGR)];
            }
        }
    }
    if (!LAST) tile_barrier<BAR_THREADS>();
}

// One length-N forward FFT down each of the CW columns of a tile.  Needs >= col_threads(N) threads in
// the barrier group; threads beyond col_threads(N) only take part in the barriers.
// smem: tile buffer (col_tile_bytes(N)); ptw: table filled by fill_pass_twiddles<N>.
template <int N, int BAR_THREADS = 0, class LD, class ST, class HOOK = NoHook, class HOOK2 = NoHook>
__device__ __forceinline__ void tile_fft(cd *smem, const cd *ptw, LD &ld, ST &st, HOOK after_load = HOOK(), HOOK2 tile_dead = HOOK2())
{
    constexpr int PT = col_pt(N);
    constexpr int NP = col_npass(N);
    constexpr int R0 = col_radix(N, 0);
    const int c = threadIdx.x % CW, u = threadIdx.x / CW;
    const bool active = threadIdx.x < col_threads(N);
    cd v[PT];
    if constexpr (NP == 1) {
        col_pass<N, PT, R0, 1, true, true, BAR_THREADS>(v, smem, ptw, ld, st, u, c, active, after_load, tile_dead);
    } else if constexpr (NP == 2) {
        constexpr int R1 = col_radix(N, 1);
        col_pass<N, PT, R0, 1, true, false, BAR_THREADS>(v, smem, ptw, ld, st, u, c, active, after_load, tile_dead);
        col_pass<N, PT, R1, R0, false, true, BAR_THREADS>(v, smem, ptw, ld, st, u, c, active, after_load, tile_dead);
    } else {
        constexpr int R1 = col_radix(N, 1), R2 = col_radix(N, 2);
        col_pass<N, PT, R0, 1, true, false, BAR_THREADS>(v, smem, ptw, ld, st, u, c, active, after_load, tile_dead);
        col_pass<N, PT, R1, R0, false, false, BAR_THREADS>(v, smem, ptw, ld, st, u, c, active, after_load, tile_dead);
        col_pass<N, PT, R2, R0 * R1, false, true, BAR_THREADS>(v, smem, ptw + R0 * R1, ld, st, u, c, active, after_load, tile_dead);
    }
}

// ---- nx <= 256: whole column in one tile -------------------------------------------------------
template <int N> __host__ __device__ constexpr size_t single_smem_bytes() { return col_tile_bytes(N) + col_tw_bytes(N); }

template <int N>
__global__ void __launch_bounds__(col_threads(N)) cols_single_kernel(InterView in, ColDst out, const cd *__restrict__ tw)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *smem = reinterpret_cast<cd *>(smem_raw);
    cd *ptw = reinterpret_cast<cd *>(smem_raw + col_tile_bytes(N));
    fill_pass_twiddles<N>(ptw, tw, 1u, (int) threadIdx.x, col_threads(N));
    if (col_npass(N) > 1) __syncthreads();
    const unsigned ct = blockIdx.x;
    unsigned ctr, row0;
    coldst_strip(out, ct, ctr, row0);
    auto ld = [&](int i, int c) -> cd { return ld_stream(inter_ptr(in, (unsigned) i, ct, (unsigned) c)); };
    auto st = [&](int k, int c, cd val) {
        const unsigned kl = ctr * CW + c;
        if (kl < out.w) st_stream(coldst_ptr(out, row0 + out.vt * (unsigned) k, kl), val);
    };
    tile_fft<N>(smem, ptw, ld, st);
}

// ---- four-step level A: FFT over x1 for fixed x2, then inter-level twiddle ----------------------
// W2[x2][k1] = w_nx^(k1*x2)
template <int N1> __host__ __device__ constexpr size_t levelA_smem_bytes() { return col_tile_bytes(N1) + col_tw_bytes(N1) + N1 * sizeof(cd); }

template <int N1>
__global__ void __launch_bounds__(col_threads(N1))
    cols_levelA_kernel(InterView in, cd *__restrict__ S, unsigned n2, const cd *__restrict__ tw, const cd *__restrict__ W2)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *smem = reinterpret_cast<cd *>(smem_raw);
    cd *ptw = reinterpret_cast<cd *>(smem_raw + col_tile_bytes(N1));
    cd *wil = ptw + col_tw_entries(N1);
    const unsigned x2 = blockIdx.x, ct = blockIdx.y;
    fill_pass_twiddles<N1>(ptw, tw, n2, (int) threadIdx.x, col_threads(N1));
    for (int i = threadIdx.x; i < N1; i += col_threads(N1)) wil[i] = ldtw(W2, x2 * (unsigned) N1 + (unsigned) i);
    __syncthreads();
    cd *Sct = S + (unsigned long long) ct * N1 * n2 * CW;
    auto ld = [&](int i, int c) -> cd { return ld_stream(inter_ptr(in, (unsigned) i * n2 + x2, ct, (unsigned) c)); };
    auto st = [&](int k1, int c, cd val) { Sct[((unsigned long long) k1 * n2 + x2) * CW + c] = cmul(val, wil[k1]); };
    tile_fft<N1>(smem, ptw, ld, st);
}

// ---- four-step level B: FFT over x2 for fixed k1, output row kx = k1 + n1*k2 ---------------------
template <int N2>
__global__ void __launch_bounds__(col_threads(N2))
    cols_levelB_kernel(const cd *__restrict__ S, ColDst out, unsigned n1, const cd *__restrict__ tw)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *smem = reinterpret_cast<cd *>(smem_raw);
    cd *ptw = reinterpret_cast<cd *>(smem_raw + col_tile_bytes(N2));
    fill_pass_twiddles<N2>(ptw, tw, n1, (int) threadIdx.x, col_threads(N2));
    if (col_npass(N2) > 1) __syncthreads();
    const unsigned k1 = blockIdx.x, ct = blockIdx.y;
    const cd *Sk = S + ((unsigned long long) ct * n1 + k1) * N2 * CW;
    auto ld = [&](int i, int c) -> cd { return Sk[(unsigned) i * CW + c]; };
    unsigned ctr, row0;
    coldst_strip(out, ct, ctr, row0);
    auto st = [&](int k2, int c, cd val) {
        const unsigned kl = ctr * CW + c;
        if (kl < out.w) st_stream(coldst_ptr(out, row0 + out.vt * (k1 + n1 * (unsigned) k2), kl), val);
    };
    tile_fft<N2>(smem, ptw, ld, st);
}

// ---- fused four-step: level A and level B in ONE persistent launch, intermediate kept in L2 -------
//
// Work is ordered by column strip (= tile of CW columns): group g holds the n2 level-A tiles of strip g
// followed by the n1 level-B tiles of strip g-LAG.  CTAs claim tiles in that order from a global
// counter.  A level-B tile waits until all level-A tiles of its strip have signalled (they were claimed
// LAG groups earlier, so in steady state the wait falls through); a level-A tile that reuses a scratch
// slot waits for the level-B tiles that last read it.  Dependencies only ever point to earlier-claimed
// tiles, which run on resident CTAs, so the scheme cannot deadlock.  The scratch is a ring of NSLOT
// strips (NSLOT * n1*n2*CW*16 B, ~24 MB for 16384 rows) that stays resident in the 126 MB L2, so the
// column pass costs one HBM read and one HBM write of the array instead of two of each.
struct FusedCtl {
    unsigned *counter; // next tile to claim
    unsigned *doneA;   // [ntiles] finished level-A tiles per strip
    unsigned *doneB;   // [ntiles] finished level-B tiles per strip
    unsigned lag, nslot;
    unsigned discard; // level-B tiles discard their scratch lines from L2 after reading them
    unsigned ct0;     // first strip of this launch (chunked column pass); counters / ring slots are launch-local
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned claim_tile(unsigned *counter)
{
    unsigned t;
    asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(counter) : "memory");
    return t;
}
// Drops a consumed 128-byte scratch line from L2 without writing it back: level-B tiles are the only
// readers of the level-A output, so once read the line is dead and its write-back would be pure waste.
__device__ __forceinline__ void l2_discard_line(const void *p)
{
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
}

template <int BAR_THREADS> __device__ __forceinline__ void cta_signal(unsigned *p)
{
    tile_barrier<BAR_THREADS>(); // all of this CTA's stores are issued (and its smem tile is free again)
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(p, 1u);
    }
}

template <int N1, int N2> __host__ __device__ constexpr int fused_threads()
{
    return col_threads(N1) > col_threads(N2) ? col_threads(N1) : col_threads(N2);
}
template <int N1, int N2> __host__ __device__ constexpr size_t fused_tile_bytes()
{
    return col_tile_bytes(N1) > col_tile_bytes(N2) ? col_tile_bytes(N1) : col_tile_bytes(N2);
}
// tile | pass twiddles N1 | pass twiddles N2 (shared with N1's when N1 == N2: same contents) | inter-level row
// (+ w_{SPLIT N1}^i, i < N1, for the radix-SPLIT pre-stage)
template <int N1, int N2, int SPLIT = 1> __host__ __device__ constexpr size_t fused_smem_bytes()
{
    return fused_tile_bytes<N1, N2>() + col_tw_bytes(N1) + (N1 == N2 ? 0 : col_tw_bytes(N2)) + N1 * sizeof(cd) +
           (SPLIT > 1 ? N1 * sizeof(cd) : 0);
}
// resident CTAs per SM the register allocator is asked to make room for
#ifndef HPXFFT_B200_FUSED_MINBLOCKS
#define HPXFFT_B200_FUSED_MINBLOCKS 5
#endif
template <int N1, int N2> __host__ __device__ constexpr int fused_min_blocks()
{
    return fused_threads<N1, N2>() <= 128 ? HPXFFT_B200_FUSED_MINBLOCKS : (fused_threads<N1, N2>() <= 256 ? 2 : 1);
}

// SPLIT = 2: columns of length nx = 2 n', n' = N1 N2.  One decimation-in-frequency stage over the leading index,
// y_c2[x'] = w_nx^(x' c2) (Y[x'] + (-1)^c2 Y[x' + n']), is folded into the level-A load, and the two length-n'
// transforms (spectrum rows kx = c2 + 2 k') run as two adjacent "virtual strips" through the same tiles.  nx = 32768 thus
// keeps the 128-point tiles (32 KB, 5 CTAs per SM) of the 16384 case instead of 256-point tiles at 2 CTAs per SM;
// the price is that every input element is loaded twice (the second time from L2).  `ntiles` counts virtual strips.
template <int N1, int N2, int SPLIT = 1>
__global__ void __launch_bounds__(fused_threads<N1, N2>(), fused_min_blocks<N1, N2>())
    cols_fused_kernel(InterView in, cd *__restrict__ S, ColDst out, const cd *__restrict__ tw, const cd *__restrict__ W2, unsigned ntiles,
                      FusedCtl ctl)
{
    static_assert(SPLIT == 1 || SPLIT == 2, "radix of the pre-stage");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NT = fused_threads<N1, N2>();
    cd *smem = reinterpret_cast<cd *>(smem_raw);
    cd *ptw1 = reinterpret_cast<cd *>(smem_raw + fused_tile_bytes<N1, N2>());
    cd *ptw2 = N1 == N2 ? ptw1 : ptw1 + col_tw_entries(N1);
    cd *wil = ptw2 + col_tw_entries(N2);
    cd *wsp = wil + N1; // SPLIT == 2 only
    __shared__ unsigned s_tile;
    constexpr unsigned PER_GROUP = N1 + N2;
    const unsigned total = (ntiles + ctl.lag) * PER_GROUP;
    const unsigned long long slot_elems = (unsigned long long) N1 * N2 * CW;

    fill_pass_twiddles<N1>(ptw1, tw, (unsigned) (SPLIT * N2), (int) threadIdx.x, NT);
    if (N1 != N2) fill_pass_twiddles<N2>(ptw2, tw, (unsigned) (SPLIT * N1), (int) threadIdx.x, NT);
    if constexpr (SPLIT > 1) {
        for (int i = threadIdx.x; i < N1; i += NT) wsp[i] = ldtw(tw, (unsigned) i * (unsigned) N2); // w_nx^(i N2) = w_{SPLIT N1}^i
    }

    // dependency of tile t: (counter address, target) -- nullptr when there is none
    auto dep_of = [&](unsigned t, unsigned &target) -> const unsigned * {
        target = 0;
        if (t >= total) return nullptr;
        const unsigned g = t / PER_GROUP, r = t - g * PER_GROUP;
        if (r < (unsigned) N2) {
            if (g >= ntiles || g < ctl.nslot) return nullptr;
            target = (unsigned) N1;
            return ctl.doneB + (g - ctl.nslot);
        }
        if (g < ctl.lag) return nullptr;
        target = (unsigned) N2;
        return ctl.doneA + (g - ctl.lag);
    };

    // Thread 0 runs a two-deep claim queue so that neither the claim atomic nor the dependency poll
    // of the next tile sits on the critical path: both are issued while the current tile is computed.
    // Two CTA barriers per tile: the one between the Stockham passes, and one at the end of the tile that
    // simultaneously (i) orders this tile's stores before its completion signal, (ii) frees the tile buffer
    // and (iii) publishes the next tile id that thread 0 wrote just before it.
    __shared__ unsigned s_ready;
    unsigned t_cur = 0, t_next = 0, dep_target = 0;
    const unsigned *dep_ptr = nullptr;
    if (threadIdx.x == 0) {
        t_cur = claim_tile(ctl.counter);
        t_next = claim_tile(ctl.counter);
        dep_ptr = dep_of(t_cur, dep_target);
        if (dep_ptr)
            while (ld_acquire_u32(dep_ptr) < dep_target) __nanosleep(64);
        s_tile = t_cur;
        s_ready = 1u;
    }
    __syncthreads();
    for (;;) {
        const unsigned t = s_tile;
        if (t >= total) break;
        unsigned t_nn = 0, next_seen = 0, next_target = 0;
        const unsigned *next_ptr = nullptr;
        if (threadIdx.x == 0) {
            t_nn = claim_tile(ctl.counter);                 // consumed at the end of this iteration
            next_ptr = dep_of(t_next, next_target);
            if (next_ptr) next_seen = ld_acquire_u32(next_ptr); // early poll of the next tile's dependency
        }
        const unsigned g = t / PER_GROUP, r = t - g * PER_GROUP;
        unsigned *done = nullptr;
        bool synced = false; // did every thread pass a CTA barrier since it read s_tile?
        if (r < (unsigned) N2) {
            // level A: tile x2 = r of strip g
            if (g < ntiles) {
#ifdef HPXFFT_B200_DIAG_WRAP
                const unsigned x2 = r, ct = (ctl.ct0 + g / SPLIT) & 1u;
#else
                const unsigned x2 = r, ct = ctl.ct0 + g / SPLIT;
#endif
                for (int i = threadIdx.x; i < N1; i += NT) wil[i] = ldtw(W2, x2 * (unsigned) N1 + (unsigned) i);
                cd *Sct = S + (unsigned long long) (g % ctl.nslot) * slot_elems;
                auto st = [&](int k1, int c, cd val) { st_cg(Sct + ((unsigned long long) k1 * N2 + x2) * CW + c, cmul(val, wil[k1])); };
                if constexpr (SPLIT == 1) {
#ifdef HPXFFT_B200_DIAG_NOLOAD_I
                    auto ld = [&](int i, int c) -> cd { return make_double2((double) (i + x2), (double) c); };
#else
                    auto ld = [&](int i, int c) -> cd { return ld_stream(inter_ptr(in, (unsigned) i * N2 + x2, ct, (unsigned) c)); };
#endif
                    tile_fft<N1>(smem, ptw1, ld, st);
                } else {
                    const unsigned c2 = g % SPLIT;
                    const cd ws = ldtw(tw, x2); // w_nx^x2
                    // both halves of the column are read with default caching: the sibling virtual strip re-reads them from L2
                    auto ld = [&](int i, int c) -> cd {
                        const unsigned x = (unsigned) i * N2 + x2;
                        const cd a = ld_cg(inter_ptr(in, x, ct, (unsigned) c)), b = ld_cg(inter_ptr(in, x + (unsigned) (N1 * N2), ct, (unsigned) c));
                        return c2 ? cmul(csub(a, b), cmul(ws, wsp[i])) : cadd(a, b);
                    };
                    tile_fft<N1>(smem, ptw1, ld, st);
                }
                done = ctl.doneA + g;
                synced = col_npass(N1) >= 2;
            }
        } else if (g >= ctl.lag) {
            // level B: tile k1 = r - N2 of strip g - lag
#ifdef HPXFFT_B200_DIAG_WRAP
            const unsigned k1 = r - N2, sl = g - ctl.lag, ct = (ctl.ct0 + sl / SPLIT) & 1u;
#else
            const unsigned k1 = r - N2, sl = g - ctl.lag, ct = ctl.ct0 + sl / SPLIT;
#endif
            const unsigned c2 = sl % SPLIT;
            const cd *Sk = S + (unsigned long long) (sl % ctl.nslot) * slot_elems + (unsigned long long) k1 * N2 * CW;
            auto ld = [&](int i, int c) -> cd { return ld_cg(Sk + (unsigned) i * CW + c); };
            unsigned ctr, row0;
            coldst_strip(out, ct, ctr, row0);
            auto st = [&](int k2, int c, cd val) {
                const unsigned kl = ctr * CW + c;
                const unsigned kx = row0 + out.vt * (c2 + (unsigned) SPLIT * (k1 + (unsigned) N1 * (unsigned) k2));
#ifdef HPXFFT_B200_DIAG_NOSTORE_V
                if (kl < out.w && val.x == 1.2345678e300) st_stream(coldst_ptr(out, kx, kl), val);
#else
                if (kl < out.w) st_stream(coldst_ptr(out, kx, kl), val);
#endif
            };
            tile_fft<N2>(smem, ptw2, ld, st);
            synced = col_npass(N2) >= 2;
            if (!synced) {
                __syncthreads(); // single-pass tile: no internal barrier, but the discard below needs one
                synced = true;
            }
            if (ctl.discard) {
                // every thread's loads of this tile have been consumed (one barrier ago): retire the lines
                for (int ln = threadIdx.x; ln < N2 * CW * (int) sizeof(cd) / 128; ln += NT)
                    l2_discard_line(reinterpret_cast<const char *>(Sk) + (size_t) ln * 128);
            }
            done = ctl.doneB + sl;
        }
        if (!synced) __syncthreads(); // skipped / single-pass tile: everybody has read s_tile before it changes
        // publish the next tile; if its dependency is not known to be satisfied yet, say so and take the
        // slow path below -- thread 0 must never spin while it still owes this tile's completion signal
        if (threadIdx.x == 0) {
            s_tile = t_next;
            s_ready = (!next_ptr || next_seen >= next_target) ? 1u : 0u;
        }
        __syncthreads();
        if (threadIdx.x == 0 && done) {
            // level A: the scratch stores of this tile must be visible before a level-B tile may read them.  level B: the counter
            // tells later level-A tiles that this tile has finished READING its ring slot (a write-after-read hazard) and, when
            // the consumed lines are discarded from L2, that the discards are done (a discard overtaken by the next write
            // would drop it).  The output stores themselves need no fence -- kernel completion publishes them -- so without
            // discards (fused transport: a fence there waits for NVLink round trips on the critical path of every tile) level B
            // signals without one.
#ifndef HPXFFT_B200_DIAG_NOFENCE
            if (r < (unsigned) N2 || ctl.discard) __threadfence();
#endif
            atomicAdd(done, 1u);
        }
        if (!s_ready) {
            if (threadIdx.x == 0)
                while (ld_acquire_u32(next_ptr) < next_target) __nanosleep(64);
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            t_cur = t_next;
            t_next = t_nn;
        }
    }
}

}  // namespace hpxfft_b200

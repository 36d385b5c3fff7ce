// kernels_cols.cuh -- forward c2c FFT along x (the strided axis) on column tiles.
//
// Replaces fft_1d_c2c_inplace on the transposed array plus both local transposes of the reference
// (core/src/shared/loop.cpp:11-15,18-25,46-53; core/src/distributed/loop.cpp:12-16,87-127):
// instead of transposing so that x becomes contiguous, a CTA owns a tile of CW = 16 adjacent ky
// columns and runs the FFT down the rows.  All global accesses are 256-byte segments, the shared
// memory tile is [point][column] with the column index fastest across threads, which makes every
// shared-memory access conflict-free without padding.
//
// nx <= 256        : one Stockham FFT per tile (cols_single_kernel).
// nx = n1*n2 > 256 : four-step.  Level A (cols_levelA_kernel): for every x2, FFT over x1 (stride n2
//                    rows), multiply by w_nx^(k1*x2), write scratch S[ct][k1][x2][c].  Level B
//                    (cols_levelB_kernel): for every k1, FFT over x2 (contiguous in S), result row is
//                    kx = k1 + n1*k2, written straight to its final position (ColDst).
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
__host__ __device__ constexpr size_t col_smem_bytes(int N) { return N <= 16 ? 0 : (size_t) N * CW * sizeof(cd); }

template <int N, int PT, int R, int NS, bool FIRST, bool LAST, class LD, class ST>
__device__ __forceinline__ void col_pass(cd (&v)[PT], cd *smem, const cd *__restrict__ tw, unsigned tws, LD &ld, ST &st,
                                         int u, int c, bool active)
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
    if (!FIRST && !LAST) __syncthreads(); // every thread has read its inputs before the in-place overwrite
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
            for (int r = 1; r < R; ++r) w[r] = cmul(w[r], ldtw(tw, (unsigned) (r * k * (N / (NS * R))) * tws));
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
    if (!LAST) __syncthreads();
}

// One length-N forward FFT down each of the CW columns of a tile.  blockDim.x >= col_threads(N);
// threads beyond col_threads(N) only take part in the barriers.
template <int N, class LD, class ST>
__device__ __forceinline__ void tile_fft(cd *smem, const cd *__restrict__ tw, unsigned tws, LD &ld, ST &st)
{
    constexpr int PT = col_pt(N);
    constexpr int NP = col_npass(N);
    constexpr int R0 = col_radix(N, 0);
    const int c = threadIdx.x % CW, u = threadIdx.x / CW;
    const bool active = threadIdx.x < col_threads(N);
    cd v[PT];
    if constexpr (NP == 1) {
        col_pass<N, PT, R0, 1, true, true>(v, smem, tw, tws, ld, st, u, c, active);
    } else if constexpr (NP == 2) {
        constexpr int R1 = col_radix(N, 1);
        col_pass<N, PT, R0, 1, true, false>(v, smem, tw, tws, ld, st, u, c, active);
        col_pass<N, PT, R1, R0, false, true>(v, smem, tw, tws, ld, st, u, c, active);
    } else {
        constexpr int R1 = col_radix(N, 1), R2 = col_radix(N, 2);
        col_pass<N, PT, R0, 1, true, false>(v, smem, tw, tws, ld, st, u, c, active);
        col_pass<N, PT, R1, R0, false, false>(v, smem, tw, tws, ld, st, u, c, active);
        col_pass<N, PT, R2, R0 * R1, false, true>(v, smem, tw, tws, ld, st, u, c, active);
    }
}

// ---- nx <= 256: whole column in one tile -------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(col_threads(N)) cols_single_kernel(InterView in, ColDst out, const cd *__restrict__ tw)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *smem = reinterpret_cast<cd *>(smem_raw);
    const unsigned ct = blockIdx.x;
    auto ld = [&](int i, int c) -> cd { return ld_stream(inter_ptr(in, (unsigned) i, ct, (unsigned) c)); };
    auto st = [&](int k, int c, cd val) {
        const unsigned kl = ct * CW + c;
        if (kl < out.w) st_stream(coldst_ptr(out, (unsigned) k, kl), val);
    };
    tile_fft<N>(smem, tw, 1u, ld, st);
}

// ---- four-step level A: FFT over x1 for fixed x2, then inter-level twiddle ----------------------
template <int N1>
__global__ void __launch_bounds__(col_threads(N1))
    cols_levelA_kernel(InterView in, cd *__restrict__ S, unsigned n2, const cd *__restrict__ tw)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *smem = reinterpret_cast<cd *>(smem_raw);
    const unsigned x2 = blockIdx.x, ct = blockIdx.y;
    cd *Sct = S + (unsigned long long) ct * N1 * n2 * CW;
    auto ld = [&](int i, int c) -> cd { return ld_stream(inter_ptr(in, (unsigned) i * n2 + x2, ct, (unsigned) c)); };
    auto st = [&](int k1, int c, cd val) {
        const cd w = ldtw(tw, (unsigned) k1 * x2); // w_nx^(k1*x2)
        Sct[((unsigned long long) k1 * n2 + x2) * CW + c] = cmul(val, w);
    };
    tile_fft<N1>(smem, tw, n2, ld, st);
}

// ---- four-step level B: FFT over x2 for fixed k1, output row kx = k1 + n1*k2 ---------------------
template <int N2>
__global__ void __launch_bounds__(col_threads(N2))
    cols_levelB_kernel(const cd *__restrict__ S, ColDst out, unsigned n1, const cd *__restrict__ tw)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *smem = reinterpret_cast<cd *>(smem_raw);
    const unsigned k1 = blockIdx.x, ct = blockIdx.y;
    const cd *Sk = S + ((unsigned long long) ct * n1 + k1) * N2 * CW;
    auto ld = [&](int i, int c) -> cd { return Sk[(unsigned) i * CW + c]; };
    auto st = [&](int k2, int c, cd val) {
        const unsigned kl = ct * CW + c;
        if (kl < out.w) st_stream(coldst_ptr(out, k1 + n1 * (unsigned) k2, kl), val);
    };
    tile_fft<N2>(smem, tw, n1, ld, st);
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
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void cta_wait_count(const unsigned *p, unsigned target)
{
    if (threadIdx.x == 0) {
        while (ld_acquire_u32(p) < target) __nanosleep(64);
    }
    __syncthreads();
}
__device__ __forceinline__ void cta_signal(unsigned *p)
{
    __syncthreads(); // all of this CTA's stores are issued (and its smem tile is free again)
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(p, 1u);
    }
}

template <int N1, int N2> __host__ __device__ constexpr int fused_threads()
{
    return col_threads(N1) > col_threads(N2) ? col_threads(N1) : col_threads(N2);
}
template <int N1, int N2> __host__ __device__ constexpr size_t fused_smem_bytes()
{
    return col_smem_bytes(N1) > col_smem_bytes(N2) ? col_smem_bytes(N1) : col_smem_bytes(N2);
}

// resident CTAs per SM the register allocator is asked to make room for
template <int N1, int N2> __host__ __device__ constexpr int fused_min_blocks()
{
    return fused_threads<N1, N2>() <= 128 ? 5 : (fused_threads<N1, N2>() <= 256 ? 2 : 1);
}

__device__ __forceinline__ unsigned claim_tile(unsigned *counter)
{
    unsigned t;
    asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(counter) : "memory");
    return t;
}

template <int N1, int N2>
__global__ void __launch_bounds__(fused_threads<N1, N2>(), fused_min_blocks<N1, N2>())
    cols_fused_kernel(InterView in, cd *__restrict__ S, ColDst out, const cd *__restrict__ tw, unsigned ntiles, FusedCtl ctl)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *smem = reinterpret_cast<cd *>(smem_raw);
    __shared__ unsigned s_tile;
    constexpr unsigned PER_GROUP = N1 + N2;
    const unsigned total = (ntiles + ctl.lag) * PER_GROUP;
    const unsigned long long slot_elems = (unsigned long long) N1 * N2 * CW;

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
    unsigned t_cur = 0, t_next = 0, dep_seen = 0, dep_target = 0;
    const unsigned *dep_ptr = nullptr;
    if (threadIdx.x == 0) {
        t_cur = claim_tile(ctl.counter);
        t_next = claim_tile(ctl.counter);
        dep_ptr = dep_of(t_cur, dep_target);
        dep_seen = 0;
    }
    for (;;) {
        unsigned t_nn = 0, next_seen = 0, next_target = 0;
        const unsigned *next_ptr = nullptr;
        if (threadIdx.x == 0) {
            if (dep_ptr && dep_seen < dep_target)
                while (ld_acquire_u32(dep_ptr) < dep_target) __nanosleep(64);
            s_tile = t_cur;
        }
        __syncthreads();
        const unsigned t = s_tile;
        if (t >= total) break;
        if (threadIdx.x == 0) {
            t_nn = claim_tile(ctl.counter);                 // consumed at the end of this iteration
            next_ptr = dep_of(t_next, next_target);
            if (next_ptr) next_seen = ld_acquire_u32(next_ptr); // early poll of the next tile's dependency
        }
        const unsigned g = t / PER_GROUP, r = t - g * PER_GROUP;
        bool did = false;
        if (r < (unsigned) N2) {
            // level A: tile x2 = r of strip g
            if (g < ntiles) {
                const unsigned x2 = r, ct = g;
                cd *Sct = S + (unsigned long long) (g % ctl.nslot) * slot_elems;
                auto ld = [&](int i, int c) -> cd { return ld_stream(inter_ptr(in, (unsigned) i * N2 + x2, ct, (unsigned) c)); };
                auto st = [&](int k1, int c, cd val) {
                    const cd w = ldtw(tw, (unsigned) k1 * x2);
                    st_cg(Sct + ((unsigned long long) k1 * N2 + x2) * CW + c, cmul(val, w));
                };
                tile_fft<N1>(smem, tw, (unsigned) N2, ld, st);
                cta_signal(ctl.doneA + g);
                did = true;
            }
        } else if (g >= ctl.lag) {
            // level B: tile k1 = r - N2 of strip g - lag
            const unsigned k1 = r - N2, ct = g - ctl.lag;
            const cd *Sk = S + (unsigned long long) (ct % ctl.nslot) * slot_elems + (unsigned long long) k1 * N2 * CW;
            auto ld = [&](int i, int c) -> cd { return ld_cg(Sk + (unsigned) i * CW + c); };
            auto st = [&](int k2, int c, cd val) {
                const unsigned kl = ct * CW + c;
                if (kl < out.w) st_stream(coldst_ptr(out, k1 + (unsigned) N1 * (unsigned) k2, kl), val);
            };
            tile_fft<N2>(smem, tw, (unsigned) N1, ld, st);
            cta_signal(ctl.doneB + ct);
            did = true;
        }
        if (!did) __syncthreads(); // keep s_tile stable until every thread has read it
        if (threadIdx.x == 0) {
            t_cur = t_next;
            t_next = t_nn;
            dep_ptr = next_ptr;
            dep_seen = next_seen;
            dep_target = next_target;
        }
    }
}

}  // namespace hpxfft_b200
